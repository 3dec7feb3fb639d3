#!/bin/bash
# the driver's default bench line, final code
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("ms_per_step", round(d["ms_per_step"],4), "value", round(d["value"]), "roof", round(d["roofline"]["frac"],4))
print("e2e", {k:v for k,v in d["e2e"].items() if k not in("note","unit","host_alloc","input_bytes_on_host_per_step","d2h_bytes_per_step")})
print("other", {m:{k:v for k,v in r.items() if k in("ms_per_step","pcie_rx_gbs_nvml","value","equals_copy_mode","chunks","equals_device_resident_pass","max_abs_diff_vs_device_resident_pass")} for m,r in d["e2e_other_handoff"].items()})
print("train", d["train"]["ms_per_step"] if d.get("train") else None, "kernels", d.get("kernels"))
PY
tail -4 gpurun_out/bench.err
