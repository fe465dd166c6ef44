cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 240 python scripts/check_halo_conv.py 2>&1 | tail -17
for tr in 2 4 6; do echo TR=$tr; TOK_HALO_TR=$tr timeout 120 python scripts/halo_one.py 256 64 56 56 64 10 2>&1 | tail -1; done
for tr in 2 4 6 8; do echo TR=$tr; TOK_HALO_TR=$tr timeout 120 python scripts/halo_one.py 32 24 128 128 24 10 2>&1 | tail -1; done
for tr in 2 4 7; do echo TR=$tr; TOK_HALO_TR=$tr timeout 120 python scripts/halo_one.py 256 128 28 28 128 10 2>&1 | tail -1; done
