cd $GRAFT_REPO_ROOT
for pr in 1 0 1 0; do
TOK_STREAM_PRIO=$pr timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-torch 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('R50 prio=$pr', round(d['ms_per_step'],3), d['e2e']['value'])"
done
for pr in 1 0; do
TOK_STREAM_PRIO=$pr timeout 600 python bench.py --workload swin_t --steps 10 --warmup 3 --skip-cpu --skip-torch 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('Swin prio=$pr', round(d['ms_per_step'],3))"
TOK_STREAM_PRIO=$pr timeout 600 python bench.py --workload hrnet_seg --steps 5 --warmup 3 --skip-cpu --skip-torch 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('HRNet prio=$pr', round(d['ms_per_step'],3))"
done
