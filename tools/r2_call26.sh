#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
cap() {
  local name=$1 idx=$2
  PIPER_B200_GRAPH=0 timeout 300 ncu --set full --import-source on --clock-control none --launch-skip $((151 + idx)) --launch-count 1 -f -o gpurun_out/c26_$name python tools/ncu_step.py 2 > gpurun_out/c26_$name.log 2>&1
  ls -la gpurun_out/c26_$name.ncu-rep | awk '{print $5, $9}'
}
cap flow_in 95
cap ffn1 5
