cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_ddp_gpu.py -m gpu -q -s > gpurun_out/pytest_ddp_r2c.log 2>&1; echo "pytest ddp exit=$?"
grep -E "passed|failed|rank [0-9]\]" gpurun_out/pytest_ddp_r2c.log | head -40
grep -E "Error|error|Traceback" gpurun_out/pytest_ddp_r2c.log | head -20
for w in resnet50 resnet18_cifar; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 2 --steps 20 --warmup 5 --workload $w > gpurun_out/bench2_${w}_r2c.json 2> gpurun_out/bench2_${w}_r2c.err; echo "bench2 $w exit=$?"
python -c "
import json,sys
d=json.load(open('gpurun_out/bench2_${w}_r2c.json'))
print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['config'])
"
tail -3 gpurun_out/bench2_${w}_r2c.err
done
timeout 300 python bench.py --workload resnet18_cifar --steps 20 --warmup 5 --skip-cpu --skip-torch > gpurun_out/bench1_cifar_r2c.json 2>gpurun_out/bench1_cifar_r2c.err; python -c "
import json
d=json.load(open('gpurun_out/bench1_cifar_r2c.json'))
print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'])
"
