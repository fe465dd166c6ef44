#!/bin/bash
# ncu --set full over two windows of one eager training step (bench.py --profile-step):
#   A: launches [skipA, skipA+cntA) of the profiled step, B: [skipB, skipB+cntB).  Reports land in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=${1:-r1a}; SA=${2:-2}; CA=${3:-16}; SB=${4:-385}; CB=${5:-25}
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    --launch-skip $SA --launch-count $CA -f -o gpurun_out/full_${R}_A python bench.py --profile-step --skip-cpu \
    > gpurun_out/ncu_full_${R}_A.log 2>&1; echo "A exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    --launch-skip $SB --launch-count $CB -f -o gpurun_out/full_${R}_B python bench.py --profile-step --skip-cpu \
    > gpurun_out/ncu_full_${R}_B.log 2>&1; echo "B exit=$?"
ls -la gpurun_out/*.ncu-rep
