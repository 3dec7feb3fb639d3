#!/bin/bash
# GPU parity suite (+ optional -k filter via PDF_PYTEST_ARGS) and smoke
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 ${PDF_PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1
echo "exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|^FAILED|^E  |^ERROR" gpurun_out/pytest_gpu.log | head -60
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -4 gpurun_out/smoke.log
