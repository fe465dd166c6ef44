cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_swin_gpu.py tests/test_heads_gpu.py tests/test_big_conv_gpu.py -m gpu -q -x 2>&1 | tail -3
PYTHONPATH=. timeout 300 python scripts/linear_shapes.py 2>&1 | tail -16 | cut -c1-100
