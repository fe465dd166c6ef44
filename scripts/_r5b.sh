cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_swin_gpu.py -m gpu -q -x 2>&1 | tail -3
PYTHONPATH=. timeout 300 python scripts/attn_one.py 256 2>&1 | head -12
