cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r2u.log 2>&1; echo "pytest exit=$?"; tail -n 4 gpurun_out/pytest_gpu_r2u.log
timeout 300 python scripts/check_halo_conv.py quick 2>&1 | tail -3
TOK_BENCH_CALLS=gpurun_out/calls_hrnet_r2u.csv timeout 600 python bench.py --workload hrnet_seg --steps 5 --warmup 3 --skip-cpu > gpurun_out/bench_hrnet_r2u.json 2> gpurun_out/bench_hrnet_r2u.err; echo "exit=$?"
TOK_DIRECT_HALO=0 timeout 600 python bench.py --workload hrnet_seg --steps 5 --warmup 3 --skip-cpu > gpurun_out/bench_hrnet_r2u0.json 2> gpurun_out/bench_hrnet_r2u0.err; echo "exit=$?"
python - <<PY
import json
for f in ('gpurun_out/bench_hrnet_r2u.json','gpurun_out/bench_hrnet_r2u0.json'):
    d=json.load(open(f))
    print(d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['final_loss'], d['roofline']['frac'], d['gpu_launches'])
PY
