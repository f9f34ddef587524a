#!/bin/bash
# round 2, GPU call K: launch list (per-kernel device time) of a short trajectory run with one re-neighbouring
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/k_launches.csv \
    python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e --no-fdm-bench --no-extras > gpurun_out/k_bench.log 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(l for l in open("gpurun_out/k_launches.csv") if l.startswith('"')))
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot = {}
order = []
for r in rows[1:]:
    name = r[ik].split("(")[0][:70]
    v = float(r[iv].replace(",", ""))
    if name not in tot: order.append(name)
    tot.setdefault(name, []).append(v)
unit = rows[1][hdr.index("Metric Unit")]
for n in order:
    print("%-72s n=%3d  total %10.3f  max %9.3f %s" % (n, len(tot[n]), sum(tot[n]), max(tot[n]), unit))
PY
