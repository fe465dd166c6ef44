cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
PYTHONPATH=. timeout 600 ncu --set full --clock-control none --import-source on -k regex:window_attn_fwd -s 3 -c 1 -o gpurun_out/attn_fwd_r5 -f python scripts/attn_one.py 256 > gpurun_out/attn_ncu.log 2>&1
tail -3 gpurun_out/attn_ncu.log
ls -la gpurun_out/*.ncu-rep
