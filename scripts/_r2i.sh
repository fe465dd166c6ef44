cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TOK_BENCH_CALLS=gpurun_out/calls_r2i.csv timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-torch > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err; echo "bench exit=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench_r2i.json')); print(d['ms_per_step'])"
wc -l gpurun_out/calls_r2i.csv
