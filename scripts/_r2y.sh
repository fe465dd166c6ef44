cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_swin_gpu.py -m gpu -q -x 2>&1 | tail -3
TOK_BENCH_CALLS=gpurun_out/calls_swin_r2y.csv timeout 600 python bench.py --workload swin_t --steps 10 --warmup 3 --skip-cpu --skip-torch > gpurun_out/bench_swin_r2y.json 2> gpurun_out/bench_swin_r2y.err; echo "exit=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_swin_r2y.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['final_loss'], d['roofline']['families'].get('window attention'))
PY
