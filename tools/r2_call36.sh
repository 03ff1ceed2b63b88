#!/bin/bash
# round 2, call 36: ncu SourceCounters captures kept as .ncu-rep (read back with ncu -i ... --page source)
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
SECS="--section SourceCounters --section WarpStateStats --section SchedulerStats"
cap() {  # name, launch index within a 151-launch step (0-based)
  local name=$1 idx=$2
  PIPER_B200_GRAPH=0 timeout 240 ncu $SECS --import-source on --clock-control none --launch-skip $((151 + idx)) --launch-count 1 -f -o gpurun_out/c36_$name python tools/ncu_step.py 2 > gpurun_out/c36_$name.log 2>&1
  ls -la gpurun_out/c36_$name.ncu-rep | awk '{print $5, $9}'
}
cap rb64k3 143
cap mrf 150
cap ffn2 6
cap att 2
