cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for g in conv dgrad gemm; do timeout 120 tests/gpu/tok_selftest $g > gpurun_out/selftest_${g}_r2l.log 2>&1; echo "selftest $g exit=$?"; tail -n 2 gpurun_out/selftest_${g}_r2l.log; done
timeout 900 python -m pytest tests/test_resnet_gpu.py tests/test_hrnet_gpu.py tests/test_engine_gpu.py tests/test_reference_goldens_gpu.py -m gpu -q > gpurun_out/pytest_gpu_r2l.log 2>&1; echo "pytest exit=$?"; tail -n 3 gpurun_out/pytest_gpu_r2l.log
for m in 1 0; do
TOK_CONV_BRES=$m timeout 120 tests/gpu/tok_selftest perf 2>&1 | grep -E "l1 |l2 " 
TOK_CONV_BRES=$m TOK_BENCH_CALLS=gpurun_out/calls_r2l_$m.csv timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-torch > gpurun_out/bench_r2l_$m.json 2> gpurun_out/bench_r2l_$m.err; echo "bench bres=$m exit=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2l_$m.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['final_loss'])
PY
tail -3 gpurun_out/bench_r2l_$m.err
done
