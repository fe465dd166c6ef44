cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_hrnet_gpu.py tests/test_backbone_goldens_gpu.py tests/test_full_size_gpu.py tests/test_n4_gpu.py tests/test_reference_goldens_gpu.py tests/test_front_door_gpu.py -m gpu -q > gpurun_out/pytest_gpu_r3j.log 2>&1; echo "pytest exit=$?"; tail -n 2 gpurun_out/pytest_gpu_r3j.log; grep FAILED gpurun_out/pytest_gpu_r3j.log
for i in 1 2; do timeout 600 python bench.py --workload hrnet_seg --steps 5 --warmup 3 --skip-cpu --skip-torch 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('HRNet', round(d['ms_per_step'],3), d['value'], d['config']['final_loss'], d['gpu_launches'])"; done
