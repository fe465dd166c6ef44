cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r3l.log 2>&1; echo "pytest exit=$?"; tail -n 2 gpurun_out/pytest_gpu_r3l.log; grep FAILED gpurun_out/pytest_gpu_r3l.log
for g in conv dgrad gemm wgrad stem elem; do timeout 120 tests/gpu/tok_selftest $g 2>&1 | tail -n 1; done
for pdl in 1 0; do
TOK_PDL=$pdl timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-torch 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('R50 pdl=$pdl', round(d['ms_per_step'],3), d['e2e']['value'], d['config']['final_loss'])"
TOK_PDL=$pdl timeout 600 python bench.py --workload hrnet_seg --steps 5 --warmup 3 --skip-cpu --skip-torch 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('HRNet pdl=$pdl', round(d['ms_per_step'],3), d['config']['final_loss'])"
TOK_PDL=$pdl timeout 600 python bench.py --workload resnet18_cifar --steps 100 --warmup 5 --skip-cpu --skip-torch 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('CIFAR pdl=$pdl', round(d['ms_per_step'],4), d['config']['final_loss'])"
TOK_PDL=$pdl timeout 600 python bench.py --workload swin_t --steps 10 --warmup 3 --skip-cpu --skip-torch 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('Swin pdl=$pdl', round(d['ms_per_step'],3), d['config']['final_loss'])"
done
