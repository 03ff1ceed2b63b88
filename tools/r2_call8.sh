#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
SECS="--section SourceCounters --section WarpStateStats --section SpeedOfLight --section LaunchStats --section Occupancy --section MemoryWorkloadAnalysis --section SchedulerStats --section InstructionStats"
cap() {  # name, launch index within a step (0-based)
  local name=$1 idx=$2
  PIPER_B200_GRAPH=0 timeout 300 ncu $SECS --import-source on --clock-control none --launch-skip $((151 + idx)) --launch-count 1 -f -o gpurun_out/c8_$name python tools/ncu_step.py 2 > gpurun_out/c8_$name.log 2>&1
  ls -la gpurun_out/c8_$name.ncu-rep | awk '{print $5, $9}'
}
cap flow_in 95
cap flow_rs 96
cap dp1x1 46
cap att 2
cap rb2k7 148
cap mrf 150
cap ln 4
