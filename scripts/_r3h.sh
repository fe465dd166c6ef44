cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for rep in 1 2; do
for cfg in "1 4" "0 4" "1 2" "0 2"; do
set -- $cfg
TOK_CONV_DEFER_STATS=$1 TOK_BN_APPLY_VEC=$2 timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-torch 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('R50 defer=$1 vec=$2', round(d['ms_per_step'],3))"
done
done
for cfg in "1 4" "0 2" "1 2"; do
set -- $cfg
TOK_CONV_DEFER_STATS=$1 TOK_BN_APPLY_VEC=$2 timeout 600 python bench.py --workload resnet18_cifar --steps 100 --warmup 5 --skip-cpu --skip-torch 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('CIFAR defer=$1 vec=$2', round(d['ms_per_step'],4))"
done
