#!/bin/bash
# quick GPU iteration: tests (PDF_PYTEST_ARGS filters) + default bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 ${PDF_PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1
echo "exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|^FAILED|^E  |^ERROR" gpurun_out/pytest_gpu.log | cut -c1-300 | head -40
timeout 900 python bench.py ${PDF_BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
try:
    j = json.load(open("gpurun_out/bench.json"))
    print("value", round(j["value"]), "ms", round(j["ms_per_step"], 3), "eager", round(j["eager_ms_per_step"], 3),
          "e2e", round(j["e2e"]["value"]), "ms", round(j["e2e"]["ms_per_step"], 2), "h2d", j["e2e"]["h2d_bytes_per_step"],
          "launches", j["gpu_launches"])
    print("stages", j["stages_ms"]); print("tflops", j["stages_tflops"]); print("hbm", j["stages_hbm"])
    print("roofline", j["roofline"]); print("sustained", j["value_sustained"])
    print("kernels", j["kernels"]); print("train", j["train"]); print("cpu", j["cpu_baseline"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench.err").read()[-3000:])
PY
tail -5 gpurun_out/bench.err
