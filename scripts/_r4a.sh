cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_swin_gpu.py tests/test_backbone_goldens_gpu.py tests/test_full_size_gpu.py -m gpu -q -x 2>&1 | tail -3
for g in 1 0; do
TOK_PATCH_EMBED_GEMM=$g timeout 600 python bench.py --workload swin_t --steps 10 --warmup 3 --skip-cpu --skip-torch 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); f=d['roofline']['families']
print('Swin gemm=$g', round(d['ms_per_step'],3), d['config']['final_loss'], {k:v['ms'] for k,v in f.items() if 'patch' in k})"
done
