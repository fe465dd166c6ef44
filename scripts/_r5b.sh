cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_swin_gpu.py -m gpu -q -x 2>&1 | tail -3
PYTHONPATH=. timeout 300 python scripts/attn_one.py 256 2>&1 | head -12
timeout 400 python bench.py --workload swin_t --skip-cpu --skip-torch > gpurun_out/b_swin_r5b.json 2> gpurun_out/b_swin_r5b.err
python - <<PY
import json
d=json.loads(open('gpurun_out/b_swin_r5b.json').read().strip().splitlines()[-1])
print('swin_t', d['ms_per_step'], d['value'], d['clocks'])
for k,v in d['roofline']['families'].items(): print(k, v['ms'], v['launches'])
PY
