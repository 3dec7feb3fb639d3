#!/bin/bash
# A/B of the software-pipelined step (PipelinedStep): decoder-stream priority 0 / -1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in 0 -1; do
  echo "== PDF_PIPE_BACK_PRIO=$v"
  env PDF_PIPE_BACK_PRIO=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-sub --sustained-seconds 0 --cpu-sample-frames 2 > gpurun_out/pipe_$v.json 2> gpurun_out/pipe_$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/pipe_$v.json").read().strip().splitlines()[-1])
print("serial ms_per_step", round(d["ms_per_step"],4), "pipelined", d["pipelined"])
PY
  tail -3 gpurun_out/pipe_$v.err
done
