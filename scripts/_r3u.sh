cd $GRAFT_REPO_ROOT
for m in 1 0; do
TOK_MASKED_ADDEND=$m TOK_BENCH_CALLS=gpurun_out/calls_r3u_$m.csv timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-torch 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); f=d['roofline']['families']
print('R50 masked=$m', round(d['ms_per_step'],3), 'dgrad', f['conv dgrad']['ms'], f['conv dgrad']['frac_of_floor'], 'bwd apply', f['bn bwd apply']['ms'], 'eager', round(d['roofline']['eager_step_ms'],2), 'frac', round(d['roofline']['frac'],4))"
done
