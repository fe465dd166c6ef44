cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_swin_gpu.py -m gpu -q -x 2>&1 | tail -5
for gd in 1 0; do
TOK_GELU_DGRAD=$gd timeout 400 python bench.py --workload swin_t --skip-cpu --skip-torch > gpurun_out/b_swin_gd$gd.json 2> gpurun_out/b_swin_gd$gd.err
python - <<PY
import json
d=json.loads(open('gpurun_out/b_swin_gd$gd.json').read().strip().splitlines()[-1])
print('swin_t gelu_dgrad=$gd', d['ms_per_step'], d['value'], d['clocks'])
for k,v in d['roofline']['families'].items(): print('   ', k, v['ms'], v['launches'])
PY
done
