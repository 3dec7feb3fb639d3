#!/bin/bash
# zero-copy pyramid hand-off: GPU tests + a quick bench line (all e2e schedules)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q -x --timeout 150 -k "host_resident or captured_step" 2>&1 | tail -5
timeout 300 python bench.py --steps 20 --warmup 3 --no-sub --sustained-seconds 0 --cpu-sample-frames 2 ${PDF_BENCH_ARGS} > gpurun_out/zc.json 2> gpurun_out/zc.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/zc.json").read().strip().splitlines()[-1])
print("ms_per_step", round(d["ms_per_step"],4))
print("e2e", {k:v for k,v in d["e2e"].items() if k not in("note","unit","host_alloc","input_bytes_on_host_per_step","d2h_bytes_per_step")})
print("other", {m:{k:v for k,v in r.items() if k in("ms_per_step","pcie_rx_gbs_nvml","value","equals_copy_mode","chunks")} for m,r in d["e2e_other_handoff"].items()})
PY
tail -4 gpurun_out/zc.err
