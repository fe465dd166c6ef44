cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_halo_conv_gpu.py -m gpu -q -x -k masked 2>&1 | tail -3
for m in 1 0; do
TOK_MASKED_ADDEND=$m timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-torch 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); f=d['roofline']['families']
print('R50 masked=$m', round(d['ms_per_step'],3), 'dgrad', f['conv dgrad']['ms'], 'bwd apply', f['bn bwd apply']['ms'], 'eager', round(d['roofline']['eager_step_ms'],2), 'frac', round(d['roofline']['frac'],4))"
done
