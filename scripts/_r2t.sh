cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TOK_BENCH_CALLS=gpurun_out/calls_hrnet_r2t.csv timeout 600 python bench.py --workload hrnet_seg --steps 5 --warmup 3 --skip-cpu > gpurun_out/bench_hrnet_r2t.json 2> gpurun_out/bench_hrnet_r2t.err; echo "exit=$?"
timeout 600 python -m pytest tests/test_hrnet_gpu.py tests/test_backbone_goldens_gpu.py tests/test_full_size_gpu.py -m gpu -q -x 2>&1 | tail -3
python - <<PY
import json
for f in ('gpurun_out/bench_hrnet_r2t.json',):
    d=json.load(open(f))
    print(d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['final_loss'], d['roofline']['frac'])
PY
