cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_r2h.log 2>&1; echo "pytest exit=$?"
grep -E "passed|failed" gpurun_out/pytest_gpu_r2h.log | tail -n 3
grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu_r2h.log | head -30
for m in 1 0; do
TOK_BN_FUSE_APPLY=$m timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-torch > gpurun_out/bench_r2h_$m.json 2> gpurun_out/bench_r2h_$m.err; echo "bench fuse_apply=$m exit=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2h_$m.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['final_loss'], d['gpu_launches']/d['steps'])
PY
tail -3 gpurun_out/bench_r2h_$m.err
done
timeout 600 python bench.py --workload swin_t --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench_swin_r2h.json 2> gpurun_out/bench_swin_r2h.err; python -c "
import json
d=json.load(open('gpurun_out/bench_swin_r2h.json')); print('swin', d['ms_per_step'], d['value'], d['gpu_launches']/d['steps'])"
timeout 600 python bench.py --workload resnet18_cifar --steps 20 --warmup 5 --skip-cpu --skip-torch > gpurun_out/bench_cifar_r2h.json 2> gpurun_out/bench_cifar_r2h.err; python -c "
import json
d=json.load(open('gpurun_out/bench_cifar_r2h.json')); print('cifar', d['ms_per_step'], d['value'], d['gpu_launches']/d['steps'])"
