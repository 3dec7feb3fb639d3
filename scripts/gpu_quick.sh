#!/bin/bash
# quick GPU iteration: selected tests + bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 ${PDF_PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1
echo "exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|^FAILED|^E  " gpurun_out/pytest_gpu.log | head -40
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
try:
    j = json.load(open("gpurun_out/bench.json"))
    print("value", round(j["value"]), "ms", round(j["ms_per_step"], 3), "e2e", round(j["e2e"]["value"]), "launches", j["gpu_launches"])
    print("stages", j["stages_ms"]); print("tflops", j["stages_tflops"]); print("roofline", j["roofline"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench.err").read()[-3000:])
PY
