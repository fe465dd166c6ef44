cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TOK_HALO_DEBUG=1 timeout 120 python scripts/halo_one.py 256 64 56 56 64 10 2>&1 | sort | uniq -c | tail -5
TOK_HALO_DEBUG=1 timeout 120 python scripts/halo_one.py 256 128 28 28 128 10 2>&1 | sort | uniq -c | tail -5
TOK_HALO_DEBUG=1 timeout 120 python scripts/halo_one.py 32 24 128 128 24 10 2>&1 | sort | uniq -c | tail -5
TOK_HALO_DEBUG=1 timeout 120 python scripts/halo_one.py 32 40 64 64 40 10 2>&1 | sort | uniq -c | tail -5
for tr in 2 4 6 8; do echo TR=$tr; TOK_HALO_TR=$tr timeout 120 python scripts/halo_one.py 256 64 56 56 64 10 2>&1 | tail -1; done
for tr in 2 4 6; do echo TR=$tr; TOK_HALO_TR=$tr timeout 120 python scripts/halo_one.py 32 24 128 128 24 10 2>&1 | tail -1; done
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:conv3x3_halo --launch-skip 2 --launch-count 2 -f -o gpurun_out/halo_r2o python scripts/halo_one.py 256 64 56 56 64 2 > gpurun_out/ncu_halo_r2o.log 2>&1; echo "ncu exit=$?"
