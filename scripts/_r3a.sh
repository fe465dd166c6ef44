cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r3a.log 2>&1; echo "pytest exit=$?"; tail -n 3 gpurun_out/pytest_gpu_r3a.log
TOK_BENCH_CALLS=gpurun_out/calls_r3a.csv timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-torch > gpurun_out/bench_r3a.json 2> gpurun_out/bench_r3a.err; echo "bench exit=$?"
timeout 600 python bench.py --workload hrnet_seg --steps 5 --warmup 3 --skip-cpu > gpurun_out/bench_hrnet_r3a.json 2> gpurun_out/bench_hrnet_r3a.err; echo "exit=$?"
TOK_CONV_PROFILE=1 timeout 200 tests/gpu/tok_selftest perf > gpurun_out/selftest_perf_profile_r2.log 2>&1
python - <<PY
import json
for f in ('bench_r3a','bench_hrnet_r3a'):
    d=json.load(open('gpurun_out/'+f+'.json'))
    print(f, d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['final_loss'], d['roofline']['frac'])
    print({k:(v['ms'],v['frac_of_floor']) for k,v in d['roofline']['families'].items() if k.startswith('bn')})
PY
