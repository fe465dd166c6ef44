cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_ddp_gpu.py -m gpu -q -s > gpurun_out/pytest_ddp_r2f.log 2>&1; echo "pytest ddp exit=$?"
grep -E "passed|failed|rank [0-9]\]" gpurun_out/pytest_ddp_r2f.log | head -60
grep -E "Error|Traceback" gpurun_out/pytest_ddp_r2f.log | head -10
N=$(nvidia-smi -L | wc -l)
for w in resnet50 resnet18_cifar retrieval; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus $N --steps 20 --warmup 5 --workload $w > gpurun_out/bench${N}_${w}_r2f.json 2> gpurun_out/bench${N}_${w}_r2f.err; echo "bench$N $w exit=$?"
python - <<PY
import json
for line in open('gpurun_out/bench${N}_${w}_r2f.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print(d['n_gpus'], round(d['ms_per_step'],3), round(d['value']), d.get('e2e',{}).get('value'), d['config'].get('cuda_graph'), d['config'].get('grad_exchange'))
PY
grep -v "OMP_NUM_THREADS\|^\*\*\*\*" gpurun_out/bench${N}_${w}_r2f.err | tail -4
done
