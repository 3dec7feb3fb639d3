#!/bin/bash
# full ncu capture of the kernels matching $1 (regex) in one decoder step, summarised on the box
# usage: scripts/gpu_ncu_one.sh <kernel-regex> <count> [skip]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/prof_one.ncu-rep
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:"$1" -s ${3:-0} -c ${2:-4} -f -o gpurun_out/prof_one env PREC=bf16x3 python scripts/decoder_profile.py > gpurun_out/ncu_one.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_one.ncu-rep --top 25 > gpurun_out/prof_one_summary.txt 2>&1
ncu -i gpurun_out/prof_one.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > gpurun_out/prof_one_raw.csv.gz
ls -la gpurun_out/prof_one.ncu-rep
sz=$(stat -c %s gpurun_out/prof_one.ncu-rep); if [ "$sz" -gt 20000000 ]; then rm -f gpurun_out/prof_one.ncu-rep; fi
tail -3 gpurun_out/ncu_one.log
