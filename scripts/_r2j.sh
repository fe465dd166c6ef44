cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TOK_WGRAD_STREAM=0 timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"conv_fwd_persist_kernel<64" -c 3 -f -o gpurun_out/conv64_r2j python bench.py --profile-step --skip-cpu > gpurun_out/ncu_conv64_r2j.log 2>&1; echo "ncu conv64 exit=$?"
TOK_EXTRA_GRAPH=0 TOK_WGRAD_STREAM=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:window_attn -s 8 -c 2 -f -o gpurun_out/attn_r2j python scripts/bench_extra.py swin 64 > gpurun_out/ncu_attn_r2j.log 2>&1; echo "ncu attn exit=$?"
