cd $GRAFT_REPO_ROOT
for v in 1 0 1 0; do
TOK_CONV_BN192=$v timeout 400 python bench.py --workload swin_t --skip-cpu --skip-torch --steps 30 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('swin_t bn192=$v', d['ms_per_step'], d['value'], d['clocks'])"
done
