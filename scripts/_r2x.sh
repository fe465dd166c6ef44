cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r2x.log 2>&1; echo "pytest exit=$?"; tail -n 4 gpurun_out/pytest_gpu_r2x.log
for c in 1 0; do
TOK_BN_FUSE_CHAIN=$c timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-torch > gpurun_out/bench_r2x_$c.json 2> gpurun_out/bench_r2x_$c.err; echo "bench chain=$c exit=$?"
TOK_BN_FUSE_CHAIN=$c timeout 600 python bench.py --workload hrnet_seg --steps 5 --warmup 3 --skip-cpu > gpurun_out/bench_hrnet_r2x_$c.json 2> gpurun_out/bench_hrnet_r2x_$c.err; echo "exit=$?"
TOK_BN_FUSE_CHAIN=$c timeout 600 python bench.py --workload resnet18_cifar --steps 50 --warmup 5 --skip-cpu --skip-torch > gpurun_out/bench_cifar_r2x_$c.json 2> gpurun_out/bench_cifar_r2x_$c.err; echo "exit=$?"
done
python - <<PY
import json
for f in ('bench_r2x_1','bench_r2x_0','bench_hrnet_r2x_1','bench_hrnet_r2x_0','bench_cifar_r2x_1','bench_cifar_r2x_0'):
    try:
        d=json.load(open('gpurun_out/'+f+'.json'))
        print(f, d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['final_loss'], d['roofline']['frac'])
    except Exception as e:
        print(f, 'ERR', e)
PY
