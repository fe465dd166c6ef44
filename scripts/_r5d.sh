cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r5m.log 2>&1; echo "pytest exit=$?"; tail -n 2 gpurun_out/pytest_gpu_r5m.log; grep FAILED gpurun_out/pytest_gpu_r5m.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_final_r5m.json 2> gpurun_out/bench_final_r5m.err; echo "bench exit=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_final_r5m.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['frac_write_aware'], d['roofline_step']['frac'], d['cpu_baseline']['value'], d['gpu_torch_baseline']['value'], d['clocks'])
PY
for wl in swin_t hrnet_seg resnet18_cifar; do
timeout 400 python bench.py --workload $wl --skip-cpu --skip-torch > gpurun_out/bench_${wl}_r5m.json 2>/dev/null
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${wl}_r5m.json').read().strip().splitlines()[-1])
print('$wl', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline_step']['frac'] if d.get('roofline_step') else None, d['clocks'])
PY
done
