#!/bin/bash
# First GPU pass: fp32 parity tests, then the tcgen05 (bf16) tests under a timeout, then benches.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== fp32 tests" 
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 -k "not bf16" > gpurun_out/pytest_fp32.log 2>&1
echo "exit $?" >> gpurun_out/pytest_fp32.log
tail -5 gpurun_out/pytest_fp32.log
echo "== bf16 tests"
timeout 600 python -m pytest tests -m gpu -q --timeout 120 -k "bf16" > gpurun_out/pytest_bf16.log 2>&1
echo "exit $?" >> gpurun_out/pytest_bf16.log
tail -30 gpurun_out/pytest_bf16.log
echo "== bench fp32"
timeout 600 python bench.py --steps 5 --warmup 3 --precision fp32 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err
tail -c 3000 gpurun_out/bench_fp32.json; tail -5 gpurun_out/bench_fp32.err
echo "== bench bf16"
timeout 600 python bench.py --steps 10 --warmup 3 --precision bf16 > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
tail -c 3000 gpurun_out/bench_bf16.json; tail -5 gpurun_out/bench_bf16.err
