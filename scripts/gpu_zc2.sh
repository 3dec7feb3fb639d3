#!/bin/bash
# 2-GPU check of the default bench line's e2e (zero-copy of all three maps at N > 1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 --no-sub --sustained-seconds 0 --cpu-sample-frames 2 --e2e-zero-copy-levels l0,l1,l2 > gpurun_out/bench_2gpu_zc.json 2> gpurun_out/bench_2gpu_zc.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_2gpu_zc.json").read().strip().splitlines()[-1])
print("n", d["n_gpus"], "value", round(d["value"]), "ms", round(d["ms_per_step"],4))
print("e2e", {k:v for k,v in d["e2e"].items() if k not in("note",)})
print("other", {m:{k:v for k,v in r.items() if k in("ms_per_step","value","equals_device_resident_pass","chunks","pcie_rx_gbs_nvml")} for m,r in d["e2e_other_handoff"].items()})
PY
tail -3 gpurun_out/bench_2gpu_zc.err
