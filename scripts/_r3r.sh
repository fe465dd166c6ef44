cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 --skip-cpu --skip-torch > gpurun_out/bench8_r3r.json 2> gpurun_out/bench8_r3r.err; echo "bench8 exit=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --workload resnet18_cifar --steps 100 --warmup 5 --skip-cpu --skip-torch > gpurun_out/bench8_cifar_r3r.json 2> gpurun_out/bench8_cifar_r3r.err; echo "bench8 cifar exit=$?"
python - <<PY
import json
for f in ('bench8_r3r','bench8_cifar_r3r'):
    try:
        d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
        print(f, d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['config'].get('cuda_graph'), d['config'].get('grad_exchange'))
    except Exception as e: print(f,'ERR',e); print(open('gpurun_out/'+f+'.err').read()[-1500:])
PY
