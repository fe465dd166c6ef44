#!/bin/bash
# One gpurun call: self-test, GPU parity tests, bench line, ncu launch list of one step, ncu --set full of the top kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=${1:-r1a}
nvidia-smi > gpurun_out/nvidia_smi_$R.txt 2>&1
bash tests/gpu/run_selftest.sh > gpurun_out/selftest_$R.log 2>&1
echo "selftest done: $(grep -c 'exit=0' gpurun_out/selftest_*.log | tr '\n' ' ')"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$R.log 2>&1; echo "pytest exit=$?"; tail -n 5 gpurun_out/pytest_gpu_$R.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; echo "bench exit=$?"; cat gpurun_out/bench_$R.json; tail -n 3 gpurun_out/bench_$R.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_$R.csv python bench.py --profile-step --skip-cpu > gpurun_out/ncu_list_$R.log 2>&1
echo "ncu list exit=$? lines=$(wc -l < gpurun_out/launches_$R.csv)"
