cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r3g.log 2>&1; echo "pytest exit=$?"; tail -n 2 gpurun_out/pytest_gpu_r3g.log; grep FAILED gpurun_out/pytest_gpu_r3g.log
TOK_BENCH_CALLS=gpurun_out/calls_r3g.csv timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-torch > gpurun_out/bench_r3g.json 2> gpurun_out/bench_r3g.err; echo "bench exit=$?"
TOK_BENCH_CALLS=gpurun_out/calls_swin_r3g.csv timeout 600 python bench.py --workload swin_t --steps 10 --warmup 3 --skip-cpu --skip-torch > gpurun_out/bench_swin_r3g.json 2> gpurun_out/bench_swin_r3g.err; echo "exit=$?"
TOK_BENCH_CALLS=gpurun_out/calls_hrnet_r3g.csv timeout 600 python bench.py --workload hrnet_seg --steps 5 --warmup 3 --skip-cpu --skip-torch > gpurun_out/bench_hrnet_r3g.json 2> gpurun_out/bench_hrnet_r3g.err; echo "exit=$?"
timeout 600 python bench.py --workload resnet18_cifar --steps 50 --warmup 5 --skip-cpu --skip-torch > gpurun_out/bench_cifar_r3g.json 2> gpurun_out/bench_cifar_r3g.err; echo "exit=$?"
python - <<PY
import json
for f in ('bench_r3g','bench_swin_r3g','bench_hrnet_r3g','bench_cifar_r3g'):
    try:
        d=json.load(open('gpurun_out/'+f+'.json'))
        print(f, d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['final_loss'], d['roofline']['frac'], d['roofline']['kernel'][:30])
    except Exception as e: print(f,'ERR',e)
PY
