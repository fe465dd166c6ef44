cd $GRAFT_REPO_ROOT
PYTHONPATH=. timeout 300 python scripts/linear_shapes.py 2>&1 | tail -16
