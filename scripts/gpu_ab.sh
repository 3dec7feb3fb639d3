#!/bin/bash
# quick A/B of one environment switch on the default bench line's stage times
# usage: scripts/gpu_ab.sh VAR "v1 v2 ..." [extra pytest -k expression]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
VAR=$1; VALS=$2; KEXPR=$3
if [ -n "$KEXPR" ]; then
  timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "$KEXPR" 2>&1 | tail -5
fi
for v in $VALS; do
  echo "== $VAR=$v"
  env $VAR=$v timeout 600 python bench.py --steps 20 --warmup 3 --no-sub --sustained-seconds 0 --cpu-sample-frames 2 > gpurun_out/ab_${VAR}_$v.json 2> gpurun_out/ab_${VAR}_$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/ab_${VAR}_$v.json").read().strip().splitlines()[-1])
print("ms_per_step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],3), "roof", round(d["roofline"]["frac"],4))
print(d["stages_ms"])
PY
  tail -2 gpurun_out/ab_${VAR}_$v.err
done
