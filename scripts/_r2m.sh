cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "TOK_CONV_BRES=1"; TOK_CONV_BRES=1 timeout 300 python scripts/check_big_conv.py 2>&1 | tail -8
for m in 1 0; do
TOK_CONV_BRES=$m timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-torch > gpurun_out/bench_r2m_$m.json 2> gpurun_out/bench_r2m_$m.err; echo "bench bres=$m exit=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2m_$m.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['final_loss'])
PY
done
