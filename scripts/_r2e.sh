cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_n4_gpu.py tests/test_swin_gpu.py tests/test_resnet_gpu.py tests/test_engine_gpu.py -m gpu -q -s > gpurun_out/pytest_gpu_r2e.log 2>&1; echo "pytest exit=$?"
grep -E "passed|failed" gpurun_out/pytest_gpu_r2e.log | tail -n 3
grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu_r2e.log | head -30
grep -E "^(OCR|Unet|  output|  df|  dfeat|swinv2_tiny)" gpurun_out/pytest_gpu_r2e.log | head -n 40
grep -E "^  grad" gpurun_out/pytest_gpu_r2e.log | sort -t' ' -k5 -r | head -12
timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-torch > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; echo "bench exit=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2e.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'])
for k,v in d['roofline']['families'].items(): print(k, v)
PY
tail -3 gpurun_out/bench_r2e.err
