cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for g in conv dgrad gemm; do timeout 120 tests/gpu/tok_selftest $g > gpurun_out/selftest_${g}_r3f.log 2>&1; echo "selftest $g exit=$?"; tail -n 1 gpurun_out/selftest_${g}_r3f.log; done
timeout 200 tests/gpu/tok_selftest perf 2>&1 | grep PERF
timeout 300 python scripts/check_big_conv.py 2>&1 | tail -8
