cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_swin_gpu.py -m gpu -q -x 2>&1 | tail -2
timeout 600 python bench.py --workload swin_t --steps 10 --warmup 3 --skip-cpu --skip-torch 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); f=d['roofline']['families']
print('Swin', round(d['ms_per_step'],3), 'LN', f['layernorm'], 'gelu', f['gelu']['ms'])"
