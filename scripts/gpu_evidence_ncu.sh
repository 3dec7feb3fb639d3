#!/bin/bash
# Round evidence, part B: ncu launch lists and full captures.  The .ncu-rep files are summarised ON the box
# (scripts/ncu_summary.py + raw csv, gzip) and deleted when large: gpurun only copies back <= 64 MiB.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
# launch list of the bench command (eager launches; warm-up 3 + 2 steps)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
   --log-file gpurun_out/launches.csv python bench.py --no-graph --no-sub --sustained-seconds 0 --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
echo "launch rows: $(wc -l < gpurun_out/launches.csv)"
# full capture of the hot-path kernels of one step, decoder glue / GEMM / attention samples
timeout 1500 ncu --set full --clock-control none --import-source on \
   -k regex:"^(sa_mlp|gemm_bf16|knn_ball|pyramid|depth2pcl|mano|split_coeff|joint_regress)" -s 75 -c 24 \
   -f -o gpurun_out/prof_step python bench.py --no-graph --no-sub --no-decoder --sustained-seconds 0 --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:"^(mha_tc|graph_cheby|row_combine|decoder_heads|gemm_bf16)" -c 16 \
   -f -o gpurun_out/prof_dec env PREC=bf16x3 python scripts/decoder_profile.py > gpurun_out/ncu_dec.log 2>&1
PREC=bf16x3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/dec_launches.csv python scripts/decoder_profile.py > gpurun_out/dec_prof.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/train_launches.csv python scripts/train_profile.py > gpurun_out/train_prof.log 2>&1
for r in prof_step prof_dec; do
  if [ -f gpurun_out/$r.ncu-rep ]; then
    python scripts/ncu_summary.py gpurun_out/$r.ncu-rep --top 8 > gpurun_out/${r}_summary.txt 2>&1
    ncu -i gpurun_out/$r.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > gpurun_out/${r}_raw.csv.gz
    sz=$(stat -c %s gpurun_out/$r.ncu-rep)
    if [ "$sz" -gt 20000000 ]; then rm -f gpurun_out/$r.ncu-rep; fi
  fi
done
ls -la gpurun_out | tail -30
du -sh gpurun_out
