cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TOK_BENCH_CALLS=gpurun_out/calls_r3q.csv timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-torch > gpurun_out/bench_r3q.json 2> gpurun_out/bench_r3q.err; echo "bench exit=$?"; tail -3 gpurun_out/bench_r3q.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r3q.json').read().strip().splitlines()[-1])
r=d['roofline']
print(d['ms_per_step'], r['frac'], r['frac_write_aware'], r['hbm_write_gbs'])
for k,v in r['families'].items(): print(k, v['launches'], v['ms'], v['frac_of_floor'], v.get('frac_of_write_aware_floor'))
PY
