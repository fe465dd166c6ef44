cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ddp_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --skip-cpu --skip-torch > gpurun_out/bench2_r3m.json 2> gpurun_out/bench2_r3m.err; echo "bench2 exit=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload resnet18_cifar --steps 100 --warmup 5 --skip-cpu --skip-torch > gpurun_out/bench2_cifar_r3m.json 2> gpurun_out/bench2_cifar_r3m.err; echo "bench2 cifar exit=$?"
python - <<PY
import json
for f in ('bench2_r3m','bench2_cifar_r3m'):
    try:
        d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
        print(f, d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['config'].get('cuda_graph'), d['config'].get('grad_exchange'))
    except Exception as e: print(f,'ERR',e); print(open('gpurun_out/'+f+'.err').read()[-1500:])
PY
