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
# CTA-pair (cta_group::2) conv kernel, opt-in: first the mechanism probe, then the pair shapes with and without it
echo "===== gemm2cta_probe"
timeout 60 tests/gpu/gemm2cta_probe > gpurun_out/gemm2cta_probe.log 2>&1; echo "exit=$?" >> gpurun_out/gemm2cta_probe.log
tail -n 3 gpurun_out/gemm2cta_probe.log
echo "===== group pair (baseline 128x256 tiles)"
TOK_CONV_BN=256 timeout 120 tests/gpu/tok_selftest pair > gpurun_out/selftest_pair_base.log 2>&1
echo "exit=$?" >> gpurun_out/selftest_pair_base.log; tail -n 12 gpurun_out/selftest_pair_base.log
echo "===== group pair (TOK_CONV_2CTA=1)"
TOK_CONV_2CTA=1 TOK_CONV_BN=256 timeout 120 tests/gpu/tok_selftest pair > gpurun_out/selftest_pair_2cta.log 2>&1
echo "exit=$?" >> gpurun_out/selftest_pair_2cta.log; tail -n 12 gpurun_out/selftest_pair_2cta.log
echo "===== perf A/B: 128x256 tiles vs CTA-pair 256x256 tiles (only shapes with M % 256 == 0, N % 256 == 0 differ)"
TOK_CONV_BN=256 timeout 120 tests/gpu/tok_selftest perf > gpurun_out/selftest_perf_bn256.log 2>&1; tail -n 14 gpurun_out/selftest_perf_bn256.log
TOK_CONV_2CTA=1 TOK_CONV_BN=256 timeout 120 tests/gpu/tok_selftest perf > gpurun_out/selftest_perf_2cta.log 2>&1; tail -n 14 gpurun_out/selftest_perf_2cta.log
echo "===== retrieval search: CTA-pair kernel (TOK_TOPK_2CTA=1) parity (bit-exact neighbour indices) and A/B at N = 262144"
TOK_TOPK_2CTA=1 timeout 300 python -m pytest tests/test_retrieval_meter.py -m gpu -q -x > gpurun_out/retrieval_pair_tests.log 2>&1; tail -n 3 gpurun_out/retrieval_pair_tests.log
timeout 120 python scripts/bench_extra.py retrieval 262144 512 1 > gpurun_out/retrieval_ab_base.log 2>&1; tail -n 1 gpurun_out/retrieval_ab_base.log
TOK_TOPK_2CTA=1 timeout 120 python scripts/bench_extra.py retrieval 262144 512 1 > gpurun_out/retrieval_ab_pair.log 2>&1; tail -n 1 gpurun_out/retrieval_ab_pair.log
for g in ${@:-gemm conv dgrad wgrad stem elem perf}; do
  echo "===== group $g"
  timeout 120 tests/gpu/tok_selftest $g > gpurun_out/selftest_$g.log 2>&1
  echo "exit=$?" >> gpurun_out/selftest_$g.log
  tail -n 60 gpurun_out/selftest_$g.log
done
