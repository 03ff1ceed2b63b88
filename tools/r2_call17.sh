#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() {
  local name=$1; shift
  env "$@" timeout -k 10 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/c19_bench_$name.json 2> gpurun_out/c19_bench_$name.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/c19_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(f"{sys.argv[1]:22s} {d['value'] / 1e6:8.1f} M samples/s  {d['ms_per_step']:7.3f} ms  stages {[round(v, 2) for v in d['stage_ms'].values()]}  batch1 {d['batch1']['latency_ms']:.2f} ms  e2e {d['e2e']['value'] / 1e6:.1f}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  env "$@" timeout 200 python tools/tap_errors.py real 40 | cut -c1-330
  env "$@" timeout 200 python tools/tap_errors.py real 113 | cut -c1-330
}
run default X=1
run chk512 PIPER_B200_V2_CHAIN_K=512
run chk1000 PIPER_B200_V2_CHAIN_K=1000
run chains1 PIPER_B200_V2_CHAINS=1
