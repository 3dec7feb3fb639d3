#!/bin/bash
# Full GPU test-suite + bench + ncu launch list + one full ncu capture of the tensor-core kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 2500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
# launch list of one bench (device-resident loop): warmup 3 + 2 steps
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
echo "launch rows: $(wc -l < gpurun_out/launches.csv)"
# full capture of the dominant kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${PDF_NCU_KERNELS:-sa_mlp_max_kernel}" -s 6 -c 2 \
   -f -o gpurun_out/prof_sa python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
