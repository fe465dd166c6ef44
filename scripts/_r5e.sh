cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_swin_gpu.py tests/test_backbone_goldens_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 400 python bench.py --workload swin_t --skip-cpu --skip-torch --steps 30 > gpurun_out/b_swin_r5k.json 2>/dev/null; python -c "
import json
d=json.loads(open('gpurun_out/b_swin_r5k.json').read().strip().splitlines()[-1]); print('swin_t', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline_step']['frac'], d['clocks']); [print('   ',k,v['ms'],v['launches']) for k,v in d['roofline']['families'].items()]"
