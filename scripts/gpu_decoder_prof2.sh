#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PREC=bf16x3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none ${PDF_NCU_EXTRA} --profile-from-start off --csv \
   --log-file gpurun_out/dec_launches.csv python scripts/decoder_profile.py > gpurun_out/dec_prof.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/dec_launches.csv")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; kn, mv, gs = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= mv: continue
    name = r[kn].split("(")[0][:60]
    t = float(r[mv].replace(",", "")) / 1000.0
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(a[1] for a in agg.values())
print("decoder launches %d, %.1f us" % (sum(a[0] for a in agg.values()), tot))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]): print("  %-60s %4d %9.1f us %5.1f%%" % (k, a[0], a[1], 100 * a[1] / tot))
PY
timeout 600 python bench.py --workload decoder --steps 20 --warmup 3 > gpurun_out/bench_decoder.json 2> gpurun_out/bench_decoder.err
python -c "
import json; j=json.load(open('gpurun_out/bench_decoder.json')); print('decoder ms', j['ms_per_step'], 'eager', j['eager_ms_per_step'], 'launches', j['gpu_launches'])"
