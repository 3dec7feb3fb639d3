#!/bin/bash
# Final round evidence: GPU test-suite, smoke, the driver's bench line, reference arm, two input-form variants.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 300 python bench.py --steps 20 --warmup 3 --pyramid fp32-nchw --masks f32 --no-sub --sustained-seconds 0 --cpu-sample-frames 2 > gpurun_out/bench_fp32inputs.json 2> gpurun_out/bench_fp32inputs.err
tail -2 gpurun_out/bench_fp32inputs.err
timeout 300 python bench.py --frames 1 --steps 20 --warmup 3 --no-sub --sustained-seconds 0 --cpu-sample-frames 1 > gpurun_out/bench_b1.json 2> gpurun_out/bench_b1.err
tail -2 gpurun_out/bench_b1.err
python - <<'PY'
import json
for f in ("bench", "bench_fp32inputs", "bench_b1"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "ms", round(d["ms_per_step"], 4), "e2e", {k: v for k, v in d["e2e"].items() if k in ("ms_per_step", "value", "pyramid_handoff", "pcie_rx_gbs_nvml", "equals_copy_mode", "h2d_bytes_per_step")},
              "other", {m: r["ms_per_step"] for m, r in d.get("e2e_other_handoff", {}).items()}, "roof", round(d["roofline"]["frac"], 4))
    except Exception as e:
        print(f, "parse failed", e)
PY
