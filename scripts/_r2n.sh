cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TOK_BENCH_CALLS=gpurun_out/calls_hrnet_r2n.csv timeout 600 python bench.py --workload hrnet_seg --steps 5 --warmup 3 --skip-cpu > gpurun_out/bench_hrnet_r2n.json 2> gpurun_out/bench_hrnet_r2n.err; echo "exit=$?"
TOK_BENCH_CALLS=gpurun_out/calls_swin_r2n.csv timeout 600 python bench.py --workload swin_t --steps 5 --warmup 3 --skip-cpu > gpurun_out/bench_swin_r2n.json 2> gpurun_out/bench_swin_r2n.err; echo "exit=$?"
wc -l gpurun_out/calls_hrnet_r2n.csv gpurun_out/calls_swin_r2n.csv
