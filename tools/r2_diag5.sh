#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
rm -f gpurun_out/d5_taps.txt
t() { local name=$1; shift; echo "== $name: $*" >> gpurun_out/d5_taps.txt; env "$@" timeout 200 python tools/tap_errors.py tiny 20 >> gpurun_out/d5_taps.txt 2>&1; env "$@" timeout 200 python tools/tap_errors.py medium 64 >> gpurun_out/d5_taps.txt 2>&1; env "$@" timeout 300 python tools/tap_errors.py medium 128 32 >> gpurun_out/d5_taps.txt 2>&1; env "$@" timeout 300 python tools/tap_errors.py real 40 >> gpurun_out/d5_taps.txt 2>&1; }
t v2_2 PIPER_B200_V2=2
t v2_2_f16_tm PIPER_B200_V2=2 PIPER_B200_V2_PREC=f16 PIPER_B200_V2_TM=1
grep -v "^$" gpurun_out/d5_taps.txt | cut -c1-330
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout -k 10 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/d5_bench_$name.json 2> gpurun_out/d5_bench_$name.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/d5_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(f"{sys.argv[1]:22s} {d['value'] / 1e6:8.1f} M samples/s  {d['ms_per_step']:7.3f} ms  stages {[round(v, 2) for v in d['stage_ms'].values()]}  batch1 {d['batch1']['latency_ms']:.2f} ms  launches {d['gpu_launches']}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run base X=1
run v2_2 PIPER_B200_V2=2
run v2_2_tm PIPER_B200_V2=2 PIPER_B200_V2_TM=1
run v2_2_f16_tm PIPER_B200_V2=2 PIPER_B200_V2_PREC=f16 PIPER_B200_V2_TM=1
run v2_2_f16_tm_graph PIPER_B200_V2=2 PIPER_B200_V2_PREC=f16 PIPER_B200_V2_TM=1 PIPER_B200_GRAPH=1
run best_guess PIPER_B200_V2=2 PIPER_B200_V2_PREC=f16 PIPER_B200_V2_TM=1 PIPER_B200_GRAPH=1 PIPER_B200_LN2=1 PIPER_B200_POST2=1 PIPER_B200_ATT2=1 PIPER_B200_MMA=31
PIPER_B200_V2=2 PIPER_B200_V2_PREC=f16 PIPER_B200_V2_TM=1 timeout -k 10 200 python tools/layer_report.py > gpurun_out/d5_layer_report_v2_f16_tm.txt 2>&1
tail -9 gpurun_out/d5_layer_report_v2_f16_tm.txt
PIPER_B200_V2=2 PIPER_B200_V2_PREC=f16 PIPER_B200_V2_TM=1 PIPER_B200_GRAPH=1 PIPER_B200_LN2=1 PIPER_B200_POST2=1 PIPER_B200_ATT2=1 PIPER_B200_MMA=31 timeout -k 10 1000 python -m pytest tests -m gpu -q --timeout 600 -x --deselect tests/test_gpu_experimental.py > gpurun_out/d5_gpu_suite.log 2>&1
tail -15 gpurun_out/d5_gpu_suite.log | cut -c1-300
