#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
P=tools/probe/tma_probe
{ for x in -4 -8 3 1 8; do timeout 60 $P 300 132 32 $x; done; timeout 60 $P 300 136 192 -4; timeout 60 $P 300 168 64 -12; } > gpurun_out/d4_tma_probe.txt 2>&1
cat gpurun_out/d4_tma_probe.txt
t() { local name=$1; shift; echo "== $name: $*" >> gpurun_out/d4_taps.txt; env "$@" timeout 200 python tools/tap_errors.py tiny 20 >> gpurun_out/d4_taps.txt 2>&1; env "$@" timeout 200 python tools/tap_errors.py medium 64 >> gpurun_out/d4_taps.txt 2>&1; env "$@" timeout 300 python tools/tap_errors.py medium 128 32 >> gpurun_out/d4_taps.txt 2>&1; env "$@" timeout 300 python tools/tap_errors.py real 40 >> gpurun_out/d4_taps.txt 2>&1; }
t v2_2_mma3 PIPER_B200_V2=2 PIPER_B200_V2_MMA3=1
t v2_2_mma3_again PIPER_B200_V2=2 PIPER_B200_V2_MMA3=1
t v2_2_mma3_f16 PIPER_B200_V2=2 PIPER_B200_V2_MMA3=1 PIPER_B200_V2_PREC=f16
t v2_2_mma3_tm PIPER_B200_V2=2 PIPER_B200_V2_MMA3=1 PIPER_B200_V2_TM=1
t v2_2_stacked_f16 PIPER_B200_V2=2 PIPER_B200_V2_PREC=f16
grep -v "^$" gpurun_out/d4_taps.txt | cut -c1-330
PIPER_B200_V2=2 PIPER_B200_V2_TM=1 PIPER_B200_V2_MMA3=1 timeout 300 python tools/conv2_check.py > gpurun_out/d4_conv2_tm.txt 2>&1; tail -2 gpurun_out/d4_conv2_tm.txt
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout -k 10 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/d4_bench_$name.json 2> gpurun_out/d4_bench_$name.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/d4_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(f"{sys.argv[1]:22s} {d['value'] / 1e6:8.1f} M samples/s  {d['ms_per_step']:7.3f} ms  stages {[round(v, 2) for v in d['stage_ms'].values()]}  batch1 {d['batch1']['latency_ms']:.2f} ms  launches {d['gpu_launches']}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run base X=1
run v2_1_mma3 PIPER_B200_V2=1 PIPER_B200_V2_MMA3=1
run v2_2_mma3 PIPER_B200_V2=2 PIPER_B200_V2_MMA3=1
run v2_2_mma3_f16 PIPER_B200_V2=2 PIPER_B200_V2_MMA3=1 PIPER_B200_V2_PREC=f16
run v2_2_mma3_f16_tm PIPER_B200_V2=2 PIPER_B200_V2_MMA3=1 PIPER_B200_V2_PREC=f16 PIPER_B200_V2_TM=1
run v2_2_mma3_f16_tm_graph PIPER_B200_V2=2 PIPER_B200_V2_MMA3=1 PIPER_B200_V2_PREC=f16 PIPER_B200_V2_TM=1 PIPER_B200_GRAPH=1
run best_guess PIPER_B200_V2=2 PIPER_B200_V2_MMA3=1 PIPER_B200_V2_PREC=f16 PIPER_B200_V2_TM=1 PIPER_B200_GRAPH=1 PIPER_B200_LN2=1 PIPER_B200_POST2=1 PIPER_B200_ATT2=1 PIPER_B200_MMA=31
