#!/bin/bash
# compute-sanitizer over the kernels touched this round (set-abstraction early-staging kernel, GEMM, grouped decoder kernels)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
K='sa_bf16 or pointnet_plus_vs_reference or gcn_decoder_tensor_core or gemm_bf16 or gemm_split or full_size_cfg3'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x --timeout 1200 -k "$K" > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/sanitize_memcheck.log
tail -6 gpurun_out/sanitize_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q -x --timeout 1200 -k "sa_bf16 or gcn_decoder_tensor_core" > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/sanitize_racecheck.log
tail -6 gpurun_out/sanitize_racecheck.log
