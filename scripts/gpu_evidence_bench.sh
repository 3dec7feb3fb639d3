#!/bin/bash
# Round evidence, part A (small outputs): GPU test-suite, smoke, the driver's bench line and its variants.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
if [ "$1" != "quick" ]; then
timeout 600 python bench.py --steps 10 --warmup 3 --precision fp32 --no-sub --sustained-seconds 0 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err
timeout 600 python bench.py --steps 20 --warmup 3 --pyramid fp32-nchw --masks f32 --no-sub --sustained-seconds 0 > gpurun_out/bench_fp32inputs.json 2> gpurun_out/bench_fp32inputs.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-decoder --no-sub --sustained-seconds 0 > gpurun_out/bench_no_decoder.json 2> gpurun_out/bench_no_decoder.err
timeout 600 python bench.py --workload decoder --steps 20 --warmup 3 > gpurun_out/bench_decoder.json 2> gpurun_out/bench_decoder.err
timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 3 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 > gpurun_out/bench_cfg5_train.json 2> gpurun_out/bench_cfg5_train.err
timeout 600 python bench.py --frames 1 --steps 20 --warmup 3 --no-sub --sustained-seconds 0 > gpurun_out/bench_b1.json 2> gpurun_out/bench_b1.err
fi
du -sh gpurun_out
