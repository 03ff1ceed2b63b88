#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
{ for cfg in "tiny 20" "medium 64" "medium 128 32" "real 40" "medium 1" "medium 200 3"; do PIPER_B200_ATT3=1 timeout -k 5 90 python tools/tap_errors.py $cfg; echo "rc=$?"; done; } > gpurun_out/c10_att3.txt 2>&1
cut -c1-330 gpurun_out/c10_att3.txt
for name in base att3; do
  if [ $name = att3 ]; then export PIPER_B200_ATT3=1; fi
  timeout -k 10 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/c10_bench_$name.json 2> gpurun_out/c10_bench_$name.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/c10_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(f"{sys.argv[1]:22s} {d['value'] / 1e6:8.1f} M samples/s  {d['ms_per_step']:7.3f} ms  stages {[round(v, 2) for v in d['stage_ms'].values()]}  batch1 {d['batch1']['latency_ms']:.2f} ms  launches {d['gpu_launches']}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
