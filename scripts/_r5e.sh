cd $GRAFT_REPO_ROOT
PYTHONPATH=. timeout 300 python scripts/linear_shapes.py 2>&1 | tail -8 | cut -c1-130
timeout 400 python bench.py --workload swin_t --skip-cpu --skip-torch 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('swin_t', d['ms_per_step'], d['value'], d['clocks'])"
timeout 400 python bench.py --workload resnet50 --skip-cpu --skip-torch 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('resnet50', d['ms_per_step'], d['value'], d['clocks'])"
