#!/bin/bash
# full ncu capture of selected kernels from one bench step:  PDF_NCU_KERNELS=regex PDF_NCU_SKIP=n PDF_NCU_COUNT=n
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"${PDF_NCU_KERNELS}" \
   -s ${PDF_NCU_SKIP:-0} -c ${PDF_NCU_COUNT:-4} -f -o gpurun_out/${PDF_NCU_OUT:-prof} \
   python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/ncu_full.log
