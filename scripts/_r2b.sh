cd /root/repo 2>/dev/null || cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_r2b.log 2>&1; echo "pytest exit=$?"
grep -E "passed|failed" gpurun_out/pytest_gpu_r2b.log | tail -n 3
grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu_r2b.log | head -20
grep -E "^(C2|C3|C4|C5|  grad|  backbone|  eval)" gpurun_out/pytest_gpu_r2b.log | head -n 60
timeout 200 python scripts/bench_extra.py retrieval 262144 512 1 2>&1 | tail -2
TOK_TOPK_2CTA=1 timeout 200 python scripts/bench_extra.py retrieval 262144 512 1 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-torch > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "bench exit=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2b.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'])
for k,v in d['roofline']['families'].items(): print(k, v)
PY
tail -3 gpurun_out/bench_r2b.err
