cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_hrnet_r3n.csv python bench.py --workload hrnet_seg --profile-step --skip-cpu > gpurun_out/ncu_hrnet_r3n.log 2>&1; echo "ncu exit=$?"
grep -c . gpurun_out/launches_hrnet_r3n.csv
