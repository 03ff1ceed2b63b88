#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() {
  local name=$1; shift
  env "$@" timeout -k 10 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/c23_bench_$name.json 2> gpurun_out/c23_bench_$name.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/c23_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(f"{sys.argv[1]:22s} {d['value'] / 1e6:8.1f} M samples/s  {d['ms_per_step']:7.3f} ms  stages {[round(v, 2) for v in d['stage_ms'].values()]}  batch1 {d['batch1']['latency_ms']:.2f} ms {[round(v, 2) for v in d['batch1']['stage_ms'].values()]}  e2e {d['e2e']['value'] / 1e6:.1f}")
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(f"gpurun_out/c23_bench_{sys.argv[1]}.err").read()[-600:])
PY
}
run variants X=1
timeout 100 python tools/tap_errors.py medium 128 1 | cut -c1-300
PIPER_B200_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 151 --launch-count 151 --csv --log-file gpurun_out/c23_launches_b1.csv python tools/ncu_step.py 2 1 > gpurun_out/c23_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/c23_launches_b1.csv")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; kn = h.index("Kernel Name"); mv = h.index("Metric Value")
agg = collections.OrderedDict(); tot = 0.0
for r in rows[hdr + 1:]:
    if len(r) <= mv: continue
    name = r[kn].split("(")[0].replace("void pb200::<unnamed>::", "")[:60]
    us = float(r[mv].replace(",", "")) / 1000.0
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us; tot += us
print(f"batch 1: total {tot:.0f} us over {sum(a[0] for a in agg.values())} launches")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us:9.1f} us {100 * us / tot:5.1f}%  n={n:3d}  avg {us / n:7.1f}  {k}")
PY
