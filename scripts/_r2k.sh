cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TOK_WGRAD_STREAM=0 timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_fwd_persist -s 2 -c 2 -f -o gpurun_out/conv64_r2k python bench.py --profile-step --skip-cpu > gpurun_out/ncu_conv64_r2k.log 2>&1; echo "ncu conv64 exit=$?"
tail -3 gpurun_out/ncu_conv64_r2k.log
