cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_swin_r3z.csv python bench.py --workload swin_t --profile-step --skip-cpu > gpurun_out/ncu_swin_r3z.log 2>&1; echo "ncu exit=$?"
