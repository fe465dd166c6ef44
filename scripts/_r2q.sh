cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 240 python scripts/check_halo_conv.py 2>&1 | tail -17
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:conv3x3_halo --launch-skip 2 --launch-count 2 -f -o gpurun_out/halo_r2q python scripts/halo_one.py 256 64 56 56 64 2 > gpurun_out/ncu_halo_r2q.log 2>&1; echo "ncu exit=$?"
