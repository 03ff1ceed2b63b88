#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/c13_gpu_suite.log 2>&1
tail -4 gpurun_out/c13_gpu_suite.log | cut -c1-300
run() {
  local name=$1; shift
  env "$@" timeout -k 10 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/c13_bench_$name.json 2> gpurun_out/c13_bench_$name.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/c13_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(f"{sys.argv[1]:22s} {d['value'] / 1e6:8.1f} M samples/s  {d['ms_per_step']:7.3f} ms  stages {[round(v, 2) for v in d['stage_ms'].values()]}  batch1 {d['batch1']['latency_ms']:.2f} ms  e2e {d['e2e']['value'] / 1e6:.1f}  launches {d['gpu_launches']}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run default X=1
run default_again X=1
PIPER_B200_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 151 --launch-count 151 --csv --log-file gpurun_out/c13_launches.csv python tools/ncu_step.py 2 > gpurun_out/c13_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/c13_launches.csv")))
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
