#!/bin/bash
# One gpurun call for a milestone: GPU parity tests, smoke, bench line (with CPU baseline), ncu launch list of one
# step, ncu --set full of the roofline kernel.  Usage: bash scripts/gpu_round.sh <tag> [skip-parts...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=${1:-r1x}; shift
SKIP=" $* "
t0=$(date +%s)
if [[ "$SKIP" != *" tests "* ]]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$R.log 2>&1; echo "pytest exit=$? ($(( $(date +%s)-t0 )) s)"; tail -n 4 gpurun_out/pytest_gpu_$R.log
fi
if [[ "$SKIP" != *" smoke "* ]]; then
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$R.log 2>&1; echo "smoke exit=$?"; tail -n 2 gpurun_out/smoke_$R.log
fi
if [[ "$SKIP" != *" bench "* ]]; then
  timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; echo "bench exit=$? ($(( $(date +%s)-t0 )) s)"; cat gpurun_out/bench_$R.json; tail -n 3 gpurun_out/bench_$R.err
fi
if [[ "$SKIP" != *" list "* ]]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches_$R.csv python bench.py --profile-step --skip-cpu > gpurun_out/ncu_list_$R.log 2>&1
  echo "ncu list exit=$? lines=$(wc -l < gpurun_out/launches_$R.csv) ($(( $(date +%s)-t0 )) s)"
fi
if [[ "$SKIP" != *" full "* ]]; then
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_fwd_persist --launch-skip 3 \
      --launch-count 1 -f -o gpurun_out/roofline_kernel_$R python scripts/roofline_kernel.py > gpurun_out/ncu_roofline_$R.log 2>&1
  echo "ncu full exit=$? ($(( $(date +%s)-t0 )) s)"
fi
