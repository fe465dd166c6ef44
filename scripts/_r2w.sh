cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_halo_conv_gpu.py -m gpu -q -x 2>&1 | tail -5
TOK_BENCH_CALLS=gpurun_out/calls_r2w.csv timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-torch > gpurun_out/bench_r2w.json 2> gpurun_out/bench_r2w.err; echo "bench exit=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2w.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['final_loss'], d['roofline']['frac'])
PY
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:conv3x3 --launch-skip 3 --launch-count 3 -f -o gpurun_out/halo_hrnet_r2w python scripts/halo_one_w.py 32 24 128 128 24 > gpurun_out/ncu_halo_r2w.log 2>&1; echo "ncu exit=$?"
