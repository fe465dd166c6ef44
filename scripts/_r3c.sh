cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_resnet_gpu.py tests/test_hrnet_gpu.py tests/test_engine_gpu.py tests/test_reference_goldens_gpu.py tests/test_backbone_goldens_gpu.py tests/test_full_size_gpu.py tests/test_n4_gpu.py -m gpu -q -x > gpurun_out/pytest_gpu_r3c.log 2>&1; echo "pytest exit=$?"; tail -n 3 gpurun_out/pytest_gpu_r3c.log
for mb in 110 60 0; do
TOK_BN_FUSED_BWD_MB=$mb TOK_BENCH_CALLS=gpurun_out/calls_r3c_$mb.csv timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-torch > gpurun_out/bench_r3c_$mb.json 2> gpurun_out/bench_r3c_$mb.err; echo "bench mb=$mb exit=$?"
done
TOK_BN_FUSED_BWD_MB=110 timeout 600 python bench.py --workload hrnet_seg --steps 5 --warmup 3 --skip-cpu > gpurun_out/bench_hrnet_r3c.json 2> gpurun_out/bench_hrnet_r3c.err; echo "exit=$?"
TOK_BN_FUSED_BWD_MB=110 timeout 600 python bench.py --workload resnet18_cifar --steps 50 --warmup 5 --skip-cpu --skip-torch > gpurun_out/bench_cifar_r3c.json 2> gpurun_out/bench_cifar_r3c.err; echo "exit=$?"
python - <<PY
import json
for f in ('bench_r3c_110','bench_r3c_60','bench_r3c_0','bench_hrnet_r3c','bench_cifar_r3c'):
    try:
        d=json.load(open('gpurun_out/'+f+'.json'))
        print(f, d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['final_loss'], d['roofline']['frac'])
        print({k:(v['launches'],v['ms'],v['frac_of_floor']) for k,v in d['roofline']['families'].items() if k.startswith('bn')})
    except Exception as e: print(f,'ERR',e)
PY
