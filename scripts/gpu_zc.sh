#!/bin/bash
# zero-copy pyramid hand-off: GPU tests + quick bench lines over the e2e variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
i=0
for v in "--e2e-streams 2 --e2e-chunks 1" "--e2e-streams 2 --e2e-chunks 2" "--e2e-streams 4 --e2e-chunks 2" "--e2e-streams 4 --e2e-chunks 4"; do
  i=$((i+1))
  echo "== $v"
  timeout 300 python bench.py --steps 20 --warmup 3 --no-sub --sustained-seconds 0 --cpu-sample-frames 2 $v > gpurun_out/zc_$i.json 2> gpurun_out/zc_$i.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/zc_$i.json").read().strip().splitlines()[-1])
print("ms_per_step", round(d["ms_per_step"],4))
print("e2e", {k:v for k,v in d["e2e"].items() if k not in("note","unit","host_alloc","input_bytes_on_host_per_step","d2h_bytes_per_step")})
print("other", {m:{k:v for k,v in r.items() if k in("ms_per_step","pcie_rx_gbs_nvml","value")} for m,r in d["e2e_other_handoff"].items()})
PY
  tail -3 gpurun_out/zc_$i.err
done
