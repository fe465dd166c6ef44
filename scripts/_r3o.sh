cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time timeout 900 python bench.py > gpurun_out/bench_default_r3o.json 2> gpurun_out/bench_default_r3o.err ) 2>&1 | grep real
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r3o.json 2> gpurun_out/bench_ref_r3o.err ) 2>&1 | grep real
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_default_r3o.json').read().strip().splitlines()[-1])
print({k:(v if not isinstance(v,dict) else '...') for k,v in d.items()})
print('e2e',d['e2e']); print('clocks',d['clocks']); print('cpu',d['cpu_baseline']); print('torch',d.get('gpu_torch_baseline'))
print('roofline',{k:v for k,v in d['roofline'].items() if k!='families'})
r=json.loads(open('gpurun_out/bench_ref_r3o.json').read().strip().splitlines()[-1]); print('ref', r['value'], r['cpu_baseline'])
PY
