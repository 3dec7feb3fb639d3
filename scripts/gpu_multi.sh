#!/bin/bash
# multi-GPU check of the default bench (driver-style torchrun launch) + reference arm line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${PDF_N:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
   bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
tail -c 2500 gpurun_out/bench_${N}gpu.json; echo; tail -5 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
j = json.load(open("gpurun_out/bench_${N}gpu.json"))
print("N", j["n_gpus"], "value", round(j["value"]), "ms", round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"]), j["e2e"]["ms_per_step"], j["e2e"]["h2d_gbs_per_rank"])
print("train", j["train"] and {k: j["train"][k] for k in ("value","ms_per_step","allreduce_bytes_per_step","allreduce_buckets")})
PY
