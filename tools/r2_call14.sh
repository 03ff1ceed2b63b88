#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() {
  local name=$1; shift
  env "$@" timeout -k 10 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/c14_bench_$name.json 2> gpurun_out/c14_bench_$name.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/c14_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(f"{sys.argv[1]:22s} {d['value'] / 1e6:8.1f} M samples/s  {d['ms_per_step']:7.3f} ms  stages {[round(v, 2) for v in d['stage_ms'].values()]}  batch1 {d['batch1']['latency_ms']:.2f} ms  e2e {d['e2e']['value'] / 1e6:.1f}  launches {d['gpu_launches']}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run conv8 X=1
timeout 200 python tools/tap_errors.py medium 128 32 | cut -c1-300
timeout -k 10 200 python tools/layer_report.py > gpurun_out/c14_layer_report.txt 2>&1
tail -9 gpurun_out/c14_layer_report.txt
