cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_r2g.log 2>&1; echo "pytest exit=$?"
grep -E "passed|failed" gpurun_out/pytest_gpu_r2g.log | tail -n 3
grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu_r2g.log | head -30
for m in 1 0; do
TOK_WGRAD_STREAM=$m timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-torch > gpurun_out/bench_r2g_$m.json 2> gpurun_out/bench_r2g_$m.err; echo "bench wgrad_stream=$m exit=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2g_$m.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['final_loss'])
PY
tail -3 gpurun_out/bench_r2g_$m.err
done
TOK_WGRAD_STREAM=1 timeout 600 python bench.py --workload swin_t --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench_swin_r2g.json 2> gpurun_out/bench_swin_r2g.err; python -c "
import json
d=json.load(open('gpurun_out/bench_swin_r2g.json')); print('swin', d['ms_per_step'], d['value'])"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
