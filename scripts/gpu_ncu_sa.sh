#!/bin/bash
# ncu --set full capture of the set-abstraction kernels (one launch each) of one hot-path step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"${PDF_NCU_K:-sa_mlp_max}" -s ${PDF_NCU_S:-6} -c ${PDF_NCU_C:-2} \
   -f -o gpurun_out/prof_sa python bench.py --no-graph --no-sub --no-decoder --sustained-seconds 0 --steps 1 --warmup 3 > gpurun_out/ncu_sa.log 2>&1
tail -3 gpurun_out/ncu_sa.log
ls -la gpurun_out/prof_sa.ncu-rep
