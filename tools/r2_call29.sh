#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout -k 5 120 python tools/tap_errors.py medium 128 32 | cut -c1-200
timeout -k 5 120 python tools/tap_errors.py real 40 | cut -c1-200
run() {
  local name=$1; shift
  env "$@" timeout -k 10 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/c29_bench_$name.json 2> gpurun_out/c29_bench_$name.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/c29_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(f"{sys.argv[1]:22s} {d['value'] / 1e6:8.1f} M samples/s  {d['ms_per_step']:7.3f} ms  stages {[round(v, 2) for v in d['stage_ms'].values()]}  batch1 {d['batch1']['latency_ms']:.2f} ms  e2e {d['e2e']['value'] / 1e6:.1f}")
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(f"gpurun_out/c29_bench_{sys.argv[1]}.err").read()[-600:])
PY
}
run wring16 X=1
run wring4 PIPER_B200_V2_WSLOTS=4
run wring8 PIPER_B200_V2_WSLOTS=8
