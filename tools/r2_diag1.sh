#!/bin/bash
# Round-2 diagnostic call: where does conv2 lose parity, does the graph / attention fix hold, and where do the persistent
# conv kernel's warps actually wait (ncu source counters on three launches of the shipped kernel).
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
PIPER_B200_V2=2 timeout 300 python tools/conv2_check.py > gpurun_out/d1_conv2.txt 2>&1; tail -3 gpurun_out/d1_conv2.txt
PIPER_B200_V2=2 PIPER_B200_V2_TM=1 timeout 300 python tools/conv2_check.py > gpurun_out/d1_conv2_tm.txt 2>&1; tail -3 gpurun_out/d1_conv2_tm.txt
PIPER_B200_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_experimental.py -m gpu -q --timeout 300 -k "graph or env4" > gpurun_out/d1_tests.log 2>&1; tail -5 gpurun_out/d1_tests.log
SECS="--section SourceCounters --section WarpStateStats --section SpeedOfLight --section LaunchStats --section Occupancy --section MemoryWorkloadAnalysis --section SchedulerStats --section InstructionStats"
cap() {  # name, launch index within a step (0-based), extra env
  local name=$1 idx=$2; shift 2
  env "$@" timeout 300 ncu $SECS --import-source on --clock-control none --launch-skip $((157 + idx)) --launch-count 1 -f -o gpurun_out/d1_$name python tools/ncu_step.py 2 > gpurun_out/d1_$name.log 2>&1
  ncu -i gpurun_out/d1_$name.ncu-rep --page source --csv > gpurun_out/d1_$name.source.csv 2>/dev/null
  ncu -i gpurun_out/d1_$name.ncu-rep --page raw --csv > gpurun_out/d1_$name.raw.csv 2>/dev/null
  ls -la gpurun_out/d1_$name.ncu-rep | awk '{print $5, $9}'
  if [ $(stat -c %s gpurun_out/d1_$name.ncu-rep) -gt 12000000 ]; then rm gpurun_out/d1_$name.ncu-rep; fi
}
cap rb3k7 155 X=1          # generator stage 3, k = 7, dilation 12 (the slowest launch, 650 us)
cap rb3k3 150 X=1          # generator stage 3, k = 3, dilation 1
cap flow_in 95 X=1         # flow in_layer (k5, 192 -> 384, tf32x3)
cap dp1x1 46 X=1           # duration-predictor 1x1 (192 -> 192): the 32 us floor
