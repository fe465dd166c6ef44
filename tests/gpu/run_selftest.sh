#!/bin/bash
# Runs every self-test group in its own process under a timeout; logs to gpurun_out/.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
# second pass of the GEMM-shaped groups with the 128x256 tile forced wherever N > 128
for g in gemm conv dgrad; do
  echo "===== group $g (TOK_CONV_BN=256)"
  TOK_CONV_BN=256 timeout 120 tests/gpu/tok_selftest $g > gpurun_out/selftest_${g}_bn256.log 2>&1
  echo "exit=$?" >> gpurun_out/selftest_${g}_bn256.log
  tail -n 3 gpurun_out/selftest_${g}_bn256.log
done
for g in ${@:-gemm conv dgrad wgrad stem elem perf}; do
  echo "===== group $g"
  timeout 120 tests/gpu/tok_selftest $g > gpurun_out/selftest_$g.log 2>&1
  echo "exit=$?" >> gpurun_out/selftest_$g.log
  tail -n 60 gpurun_out/selftest_$g.log
done
