cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_big_conv_gpu.py tests/test_resnet_gpu.py tests/test_swin_gpu.py tests/test_halo_conv_gpu.py -m gpu -q -x 2>&1 | tail -3
for wl in swin_t resnet50 hrnet_seg; do
for mg in -1 0; do
  if [ $mg = 0 ]; then export TOK_CONV_MGROUP=0; else unset TOK_CONV_MGROUP; fi
  timeout 400 python bench.py --workload $wl --skip-cpu --skip-torch > gpurun_out/b_${wl}_mg${mg}.json 2> gpurun_out/b_${wl}_mg${mg}.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/b_${wl}_mg${mg}.json').read().strip().splitlines()[-1])
print('$wl', 'mg=$mg', d['ms_per_step'], d['value'], d['clocks'])
PY
done
done
