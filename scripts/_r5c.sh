cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
PYTHONPATH=. timeout 600 ncu --set full --clock-control none --import-source on -k regex:window_attn_bwd -s 2 -c 1 -o gpurun_out/attn_bwd_r5 -f python scripts/attn_one.py 256 > gpurun_out/attn_ncu.log 2>&1
tail -2 gpurun_out/attn_ncu.log
