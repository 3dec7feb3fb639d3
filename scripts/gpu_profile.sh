#!/bin/bash
# Round profile: full GPU test-suite, smoke, bench (both precisions), ncu launch list of the bench command,
# and one `ncu --set full` capture of every kernel of one step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --precision fp32 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 600 python bench.py --steps 20 --warmup 3 --pyramid-layout nhwc > gpurun_out/bench_nhwc.json 2> gpurun_out/bench_nhwc.err
timeout 600 python bench.py --steps 10 --warmup 3 --with-decoder > gpurun_out/bench_with_decoder.json 2> gpurun_out/bench_with_decoder.err
timeout 600 python bench.py --workload decoder --steps 20 --warmup 3 > gpurun_out/bench_decoder.json 2> gpurun_out/bench_decoder.err
timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 3 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 > gpurun_out/bench_cfg5_train.json 2> gpurun_out/bench_cfg5_train.err
timeout 600 python bench.py --workload cfg5 --precision bf16 --steps 5 --warmup 3 > gpurun_out/bench_cfg5_train_bf16.json 2> gpurun_out/bench_cfg5_train_bf16.err
timeout 900 python bench.py --workload cfg5 --impl reference --steps 1 --warmup 0 --cpu-sample-frames 2 > gpurun_out/bench_cfg5_reference.json 2> gpurun_out/bench_cfg5_reference.err
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/smi.txt 2>&1
# launch list of the bench command (warm-up 3 + 2 steps)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
   --log-file gpurun_out/launches.csv python bench.py --no-graph --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
echo "launch rows: $(wc -l < gpurun_out/launches.csv)"
# full capture of one step of our kernels (4th step)
timeout 1200 ncu --set full --clock-control none --import-source on \
   -k regex:"^(sa_mlp|gemm_bf16|knn_ball|pyramid|rows_to|depth2pcl|mano|linear_f32|split_coeff)" -s 96 -c 32 \
   -f -o gpurun_out/prof_step python bench.py --no-graph --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
# launch lists of one training step and one decoder step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/train_launches.csv python scripts/train_profile.py > gpurun_out/train_prof.log 2>&1
PREC=bf16x3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/dec_launches.csv python scripts/decoder_profile.py > gpurun_out/dec_prof.log 2>&1
ls -la gpurun_out/*.ncu-rep
