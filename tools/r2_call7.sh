#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/c7_gpu_suite.log 2>&1
tail -12 gpurun_out/c7_gpu_suite.log | cut -c1-300
timeout 600 python __graft_entry__.py smoke > gpurun_out/c7_smoke.log 2>&1; tail -3 gpurun_out/c7_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err; cut -c1-700 gpurun_out/c7_bench.json
PIPER_B200_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 151 --launch-count 151 --csv --log-file gpurun_out/c7_launches.csv python tools/ncu_step.py 2 > gpurun_out/c7_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/c7_launches.csv")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; kn = h.index("Kernel Name"); mv = h.index("Metric Value")
agg = collections.OrderedDict(); tot = 0.0
for r in rows[hdr + 1:]:
    if len(r) <= mv: continue
    name = r[kn].split("(")[0].replace("void pb200::<unnamed>::", "")[:60]
    us = float(r[mv].replace(",", "")) / 1000.0
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us; tot += us
print(f"total {tot:.0f} us over {sum(a[0] for a in agg.values())} launches")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us:9.1f} us {100 * us / tot:5.1f}%  n={n:3d}  avg {us / n:7.1f}  {k}")
PY
