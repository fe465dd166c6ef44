cd $GRAFT_REPO_ROOT
for d in 0 1; do echo DBG=$d; TOK_HALO_DBG=$d timeout 120 python scripts/halo_one.py 256 64 56 56 64 10 2>&1 | tail -1; TOK_HALO_DBG=$d timeout 120 python scripts/halo_one.py 256 128 28 28 128 10 2>&1 | tail -1; TOK_HALO_DBG=$d timeout 120 python scripts/halo_one.py 32 24 128 128 24 10 2>&1 | tail -1; done
