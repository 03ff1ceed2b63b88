#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout -k 10 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "fused or stage_by_stage or golden or ragged or multi_speaker or int16" > gpurun_out/c15_tests.log 2>&1
tail -4 gpurun_out/c15_tests.log | cut -c1-300
run() {
  local name=$1; shift
  env "$@" timeout -k 10 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/c15_bench_$name.json 2> gpurun_out/c15_bench_$name.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/c15_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(f"{sys.argv[1]:22s} {d['value'] / 1e6:8.1f} M samples/s  {d['ms_per_step']:7.3f} ms  stages {[round(v, 2) for v in d['stage_ms'].values()]}  batch1 {d['batch1']['latency_ms']:.2f} ms  e2e {d['e2e']['value'] / 1e6:.1f}  launches {d['gpu_launches']}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run mrf16_ln24 X=1
run mrf16_ln24_again X=1
