#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
P=tools/probe/tma_probe
{ for v in 300 301 310 311 200 201 210; do timeout 60 $P $v 132 32 -5; done
  timeout 60 $P 300 128 32 0; timeout 60 $P 300 64 8 0; timeout 60 $P 300 132 32 0; timeout 60 $P 300 132 32 4; timeout 60 $P 200 132 32 0; timeout 60 $P 300 136 192 -3
} > gpurun_out/d3_tma_probe.txt 2>&1
cat gpurun_out/d3_tma_probe.txt
{ echo "== twice"; PIPER_B200_V2=2 PIPER_B200_MMA=2 timeout 100 python tools/tap_errors.py tiny 20; PIPER_B200_V2=2 PIPER_B200_MMA=2 timeout 100 python tools/tap_errors.py tiny 20
  echo "== launch blocking"; CUDA_LAUNCH_BLOCKING=1 PIPER_B200_V2=2 PIPER_B200_MMA=2 timeout 100 python tools/tap_errors.py tiny 20
  echo "== f16"; PIPER_B200_V2_PREC=f16 PIPER_B200_V2=2 PIPER_B200_MMA=2 timeout 100 python tools/tap_errors.py tiny 20
} > gpurun_out/d3_v2.txt 2>&1
cut -c1-300 gpurun_out/d3_v2.txt
for tool in memcheck racecheck initcheck; do
  echo "== $tool" >> gpurun_out/d3_sanitizer.txt
  PIPER_B200_V2=2 PIPER_B200_MMA=2 timeout 400 compute-sanitizer --tool $tool python tools/tap_errors.py tiny 20 >> gpurun_out/d3_sanitizer.txt 2>&1
done
grep -E "^== |ERROR SUMMARY|RACECHECK SUMMARY|Error:|hazard|Uninitialized|Invalid|tiny/20" gpurun_out/d3_sanitizer.txt | cut -c1-250 | head -60
