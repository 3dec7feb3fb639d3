#!/bin/bash
# Round-2 evidence: GPU test-suite, smoke, the driver's bench line (+ variants), ncu launch list and full captures.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 600 python bench.py --steps 10 --warmup 3 --precision fp32 --no-sub --sustained-seconds 0 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err
timeout 600 python bench.py --steps 20 --warmup 3 --pyramid fp32-nchw --masks f32 --no-sub --sustained-seconds 0 > gpurun_out/bench_fp32inputs.json 2> gpurun_out/bench_fp32inputs.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-decoder --no-sub --sustained-seconds 0 > gpurun_out/bench_no_decoder.json 2> gpurun_out/bench_no_decoder.err
timeout 600 python bench.py --workload decoder --steps 20 --warmup 3 > gpurun_out/bench_decoder.json 2> gpurun_out/bench_decoder.err
timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 3 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 > gpurun_out/bench_cfg5_train.json 2> gpurun_out/bench_cfg5_train.err
timeout 600 python bench.py --frames 1 --steps 20 --warmup 3 --no-sub --sustained-seconds 0 > gpurun_out/bench_b1.json 2> gpurun_out/bench_b1.err
# launch list of the bench command (eager launches; warm-up 3 + 2 steps)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
   --log-file gpurun_out/launches.csv python bench.py --no-graph --no-sub --sustained-seconds 0 --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
echo "launch rows: $(wc -l < gpurun_out/launches.csv)"
# full capture of the hot-path kernels of one step (the 4th), decoder glue / GEMM / attention samples
timeout 1500 ncu --set full --clock-control none --import-source on \
   -k regex:"^(sa_mlp|gemm_bf16|knn_ball|pyramid|depth2pcl|mano|split_coeff|joint_regress)" -s 75 -c 24 \
   -f -o gpurun_out/prof_step python bench.py --no-graph --no-sub --no-decoder --sustained-seconds 0 --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:"^(mha_tc|graph_cheby|row_combine|decoder_heads|gemm_bf16)" -c 40 \
   -f -o gpurun_out/prof_dec env PREC=bf16x3 python scripts/decoder_profile.py > gpurun_out/ncu_dec.log 2>&1
PREC=bf16x3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/dec_launches.csv python scripts/decoder_profile.py > gpurun_out/dec_prof.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/train_launches.csv python scripts/train_profile.py > gpurun_out/train_prof.log 2>&1
ls -la gpurun_out/*.ncu-rep
