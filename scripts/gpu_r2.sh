#!/bin/bash
# One gpurun call of round 2: bash scripts/gpu_r2.sh <tag> part...   (parts: tests bench workloads ncu_retrieval ncu_attn list)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=${1:-r2a}; shift
PARTS=" ${*:-tests bench} "
t0=$(date +%s)
el() { echo "$(( $(date +%s)-t0 )) s"; }
if [[ "$PARTS" == *" tests "* ]]; then
  timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_$R.log 2>&1; echo "pytest exit=$? ($(el))"
  grep -E "passed|failed|error" gpurun_out/pytest_gpu_$R.log | tail -n 3
  grep -E "^(C2|C3|C4|C5|  grad|  backbone|  eval)" gpurun_out/pytest_gpu_$R.log | head -n 60
fi
if [[ "$PARTS" == *" smoke "* ]]; then
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$R.log 2>&1; echo "smoke exit=$?"; tail -n 2 gpurun_out/smoke_$R.log
fi
if [[ "$PARTS" == *" bench "* ]]; then
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; echo "bench exit=$? ($(el))"
  cat gpurun_out/bench_$R.json; tail -n 5 gpurun_out/bench_$R.err
fi
if [[ "$PARTS" == *" workloads "* ]]; then
  for w in resnet18_cifar swin_t hrnet_seg; do
    timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench_${w}_$R.json 2> gpurun_out/bench_${w}_$R.err
    echo "bench $w exit=$? ($(el))"; cat gpurun_out/bench_${w}_$R.json; tail -n 3 gpurun_out/bench_${w}_$R.err
  done
  timeout 300 python bench.py --workload retrieval --steps 2 > gpurun_out/bench_retrieval_$R.json 2> gpurun_out/bench_retrieval_$R.err
  echo "bench retrieval exit=$? ($(el))"; cat gpurun_out/bench_retrieval_$R.json; tail -n 3 gpurun_out/bench_retrieval_$R.err
fi
if [[ "$PARTS" == *" ncu_retrieval "* ]]; then
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:cosine_topk -s 1 -c 1 -f \
      -o gpurun_out/retrieval_$R python scripts/bench_extra.py retrieval 131072 512 1 > gpurun_out/ncu_retrieval_$R.log 2>&1
  echo "ncu retrieval exit=$? ($(el))"
fi
if [[ "$PARTS" == *" ncu_attn "* ]]; then
  TOK_EXTRA_GRAPH=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:window_attn -s 8 -c 2 -f \
      -o gpurun_out/attn_$R python scripts/bench_extra.py swin 64 > gpurun_out/ncu_attn_$R.log 2>&1
  echo "ncu attn exit=$? ($(el))"
fi
if [[ "$PARTS" == *" list "* ]]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches_$R.csv python bench.py --profile-step --skip-cpu > gpurun_out/ncu_list_$R.log 2>&1
  echo "ncu list exit=$? lines=$(wc -l < gpurun_out/launches_$R.csv) ($(el))"
fi
