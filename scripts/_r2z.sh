cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_swin_gpu.py -m gpu -q -x 2>&1 | tail -3
for c in 3 2; do echo CTAS=$c; TOK_ATTN_FWD_CTAS=$c timeout 100 python scripts/attn_bench.py 1 256 2>&1 | tail -2;  TOK_ATTN_FWD_CTAS=$c timeout 100 python scripts/attn_bench.py 3 256 2>&1 | tail -2; done
