#!/bin/bash
# multi-GPU check of the default bench (driver-style torchrun launch): N ranks, default line (with the cfg2 / cfg5
# sub-records) and, for the e2e staging question, the same run with torch's pin_memory() buffers (no sub-records)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${PDF_N:-2}
nvidia-smi topo -m > gpurun_out/topo_${N}gpu.txt 2>&1
lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" >> gpurun_out/topo_${N}gpu.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
   bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
tail -5 gpurun_out/bench_${N}gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
   bench.py --gpus $N --steps 20 --warmup 3 --no-sub --sustained-seconds 0 --host-alloc pinned > gpurun_out/bench_${N}gpu_pinned.json 2> gpurun_out/bench_${N}gpu_pinned.err
tail -5 gpurun_out/bench_${N}gpu_pinned.err
python - <<PY
import json
for f in ("gpurun_out/bench_${N}gpu.json", "gpurun_out/bench_${N}gpu_pinned.json"):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "N", j["n_gpus"], "value", round(j["value"]), "ms", round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"]), round(j["e2e"]["ms_per_step"],3), j["e2e"]["h2d_gbs_per_rank"], j["e2e"].get("host_alloc"))
    t = j.get("train")
    print("  cfg4", j.get("cfg4")); print("  train", t and {k: t.get(k) for k in ("value","ms_per_step","allreduce_bytes_per_step","allreduce_buckets","error")})
PY
cat gpurun_out/topo_${N}gpu.txt | head -30
