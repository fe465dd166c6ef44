cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 140 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_swin_r5.csv python bench.py --workload swin_t --steps 1 --warmup 1 --skip-cpu --skip-torch --no-graph > gpurun_out/launches_swin_r5.log 2>&1
echo "rc=$?"; wc -l gpurun_out/launches_swin_r5.csv
