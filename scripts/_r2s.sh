cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r2s.log 2>&1; echo "pytest exit=$?"; tail -n 4 gpurun_out/pytest_gpu_r2s.log
TOK_BENCH_CALLS=gpurun_out/calls_r2s.csv timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-torch > gpurun_out/bench_r2s.json 2> gpurun_out/bench_r2s.err; echo "bench exit=$?"
TOK_BENCH_CALLS=gpurun_out/calls_hrnet_r2s.csv timeout 600 python bench.py --workload hrnet_seg --steps 5 --warmup 3 --skip-cpu > gpurun_out/bench_hrnet_r2s.json 2> gpurun_out/bench_hrnet_r2s.err; echo "exit=$?"
python - <<PY
import json
for f in ('gpurun_out/bench_r2s.json','gpurun_out/bench_hrnet_r2s.json'):
    d=json.load(open(f))
    print(d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['final_loss'], d['roofline']['frac'])
PY
