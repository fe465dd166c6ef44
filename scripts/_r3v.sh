cd $GRAFT_REPO_ROOT
for g in conv dgrad gemm; do timeout 120 tests/gpu/tok_selftest $g 2>&1 | tail -n 1; done
timeout 200 tests/gpu/tok_selftest perf 2>&1 | grep -E "PERF"
timeout 300 python -m pytest tests/test_halo_conv_gpu.py tests/test_resnet_gpu.py -m gpu -q -x 2>&1 | tail -2
for m in 1 0; do
TOK_MASKED_ADDEND=$m TOK_BENCH_CALLS=gpurun_out/calls_r3v_$m.csv timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-torch 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); f=d['roofline']['families']
print('R50 masked=$m', round(d['ms_per_step'],3), 'dgrad', f['conv dgrad']['ms'], f['conv dgrad']['frac_of_floor'], 'bwd apply', f['bn bwd apply']['ms'], 'frac', round(d['roofline']['frac'],4))"
done
