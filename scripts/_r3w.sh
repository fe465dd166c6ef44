cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r3w.log 2>&1; echo "pytest exit=$?"; tail -n 2 gpurun_out/pytest_gpu_r3w.log; grep FAILED gpurun_out/pytest_gpu_r3w.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for w in hrnet_seg swin_t resnet18_cifar; do timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --skip-cpu --skip-torch 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', round(d['ms_per_step'],4), round(d['value'],1), d['config']['final_loss'])"; done
