cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r3s.log 2>&1; echo "pytest exit=$?"; tail -n 2 gpurun_out/pytest_gpu_r3s.log; grep FAILED gpurun_out/pytest_gpu_r3s.log
for m in 1 0 1 0; do
TOK_MASKED_ADDEND=$m timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-torch 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('R50 masked=$m', round(d['ms_per_step'],3), d['config']['final_loss'], d['roofline']['families']['bn bwd apply'])"
done
for m in 1 0; do
TOK_MASKED_ADDEND=$m timeout 600 python bench.py --workload hrnet_seg --steps 5 --warmup 3 --skip-cpu --skip-torch 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('HRNet masked=$m', round(d['ms_per_step'],3), d['config']['final_loss'])"
TOK_MASKED_ADDEND=$m timeout 600 python bench.py --workload resnet18_cifar --steps 100 --warmup 5 --skip-cpu --skip-torch 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('CIFAR masked=$m', round(d['ms_per_step'],4), d['config']['final_loss'])"
done
