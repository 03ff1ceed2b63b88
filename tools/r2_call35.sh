#!/bin/bash
# round 2, call 35: per-source-line stall samples (ncu SourceCounters) of the launches that carry the step, tcgen05.mma pacing probe
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 60 tools/probe/mma_probe 64 0 8 > gpurun_out/c35_probe64.txt 2>&1; echo "rc=$?"
timeout 60 tools/probe/mma_probe 128 0 8 > gpurun_out/c35_probe128.txt 2>&1; echo "rc=$?"
timeout 60 tools/probe/mma_probe 32 0 8 > gpurun_out/c35_probe32.txt 2>&1; echo "rc=$?"
SECS="--section SourceCounters --section WarpStateStats --section SchedulerStats"
cap() {  # name, launch index within a 151-launch step (0-based)
  local name=$1 idx=$2
  PIPER_B200_GRAPH=0 timeout 240 ncu $SECS --import-source on --clock-control none --launch-skip $((151 + idx)) --launch-count 1 -f -o gpurun_out/c35_$name python tools/ncu_step.py 2 > gpurun_out/c35_$name.log 2>&1
  ncu -i gpurun_out/c35_$name.ncu-rep --page source --print-source cuda --csv 2>/dev/null | gzip > gpurun_out/c35_$name.source.csv.gz
  ncu -i gpurun_out/c35_$name.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/c35_$name.raw.csv.gz
  ls -la gpurun_out/c35_$name.ncu-rep | awk '{print $5, $9}'
  rm -f gpurun_out/c35_$name.ncu-rep
}
cap rb64k3 143
cap mrf 150
cap ffn2 6
cap att 2
cap flow_in 95
zcat gpurun_out/c35_rb64k3.source.csv.gz | head -3 | cut -c1-600
