cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_r2d.log 2>&1; echo "pytest exit=$?"
grep -E "passed|failed" gpurun_out/pytest_gpu_r2d.log | tail -n 3
grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu_r2d.log | head -30
grep -E "^(C2|C3|C4|C5|OCR|Unet|  df|  worst)" gpurun_out/pytest_gpu_r2d.log | head -n 60
timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-torch > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; echo "bench exit=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2d.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'])
for k,v in d['roofline']['families'].items(): print(k, v)
PY
tail -3 gpurun_out/bench_r2d.err
for w in swin_t hrnet_seg; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench_${w}_r2d.json 2> gpurun_out/bench_${w}_r2d.err
  echo "bench $w exit=$?"; python -c "
import json
d=json.load(open('gpurun_out/bench_${w}_r2d.json'))
print(d['ms_per_step'], d['value'], d['gpu_launches']/d['steps'])
for k,v in d['roofline']['families'].items(): print(' ', k, v)
"; tail -n 3 gpurun_out/bench_${w}_r2d.err
done
timeout 300 python bench.py --workload retrieval --steps 2 > gpurun_out/bench_retrieval_r2d.json 2> gpurun_out/bench_retrieval_r2d.err; cat gpurun_out/bench_retrieval_r2d.json; tail -3 gpurun_out/bench_retrieval_r2d.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bn_bwd_reduce2 -s 30 -c 4 -f -o gpurun_out/bnreduce_r2d python bench.py --profile-step --skip-cpu > gpurun_out/ncu_bnreduce_r2d.log 2>&1; echo "ncu exit=$?"
