cd $GRAFT_REPO_ROOT
timeout 400 python bench.py --workload swin_t --skip-cpu --skip-torch > gpurun_out/b_swin_r5f.json 2> gpurun_out/b_swin_r5f.err
python - <<PY
import json
d=json.loads(open('gpurun_out/b_swin_r5f.json').read().strip().splitlines()[-1])
print('swin_t', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline_step']['frac'], d['clocks'])
for k,v in d['roofline']['families'].items(): print('   ', k, v['ms'], v['launches'], v.get('frac_of_floor'))
PY
