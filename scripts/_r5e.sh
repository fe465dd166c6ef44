cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_resnet_gpu.py tests/test_pooling_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 400 python bench.py --workload resnet50 --skip-cpu --skip-torch --steps 30 > gpurun_out/b_r50_r5l.json 2>/dev/null; python -c "
import json
d=json.loads(open('gpurun_out/b_r50_r5l.json').read().strip().splitlines()[-1]); print('resnet50', d['ms_per_step'], d['value'], d['clocks']); [print('   ',k,v['ms'],v['launches']) for k,v in d['roofline']['families'].items() if 'stem' in k]"
