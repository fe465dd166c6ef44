cd $GRAFT_REPO_ROOT
for g in conv dgrad gemm; do timeout 120 tests/gpu/tok_selftest $g 2>&1 | tail -n 1; done
for c in 1 0; do echo CBUFS3=$c; TOK_CONV_CBUFS3=$c timeout 200 tests/gpu/tok_selftest perf 2>&1 | grep -E "PERF (l1 1x1|l3 1x1|l4 1x1|ds)"; done
timeout 300 python scripts/check_big_conv.py 2>&1 | tail -3
