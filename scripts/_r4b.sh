cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_swin_gpu.py tests/test_backbone_goldens_gpu.py tests/test_full_size_gpu.py tests/test_engine_gpu.py -m gpu -q 2>&1 | tail -3
for i in 1 2; do timeout 600 python bench.py --workload swin_t --steps 10 --warmup 3 --skip-cpu --skip-torch 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('Swin', round(d['ms_per_step'],3), round(d['value'],1), d['config']['final_loss'])"; done
