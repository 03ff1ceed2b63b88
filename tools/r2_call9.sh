#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
{ echo "== att3 tiny"; PIPER_B200_ATT3=1 timeout -k 5 40 python tools/tap_errors.py tiny 20; echo "rc=$?"
  echo "== att3 tiny under memcheck"; PIPER_B200_ATT3=1 timeout -k 5 150 compute-sanitizer --tool memcheck python tools/tap_errors.py tiny 20 2>&1 | grep -v "^=========     \|^=========         " | head -40; echo "rc=$?"
  echo "== att3 medium 64"; PIPER_B200_ATT3=1 timeout -k 5 40 python tools/tap_errors.py medium 64; echo "rc=$?"
  echo "== att3 medium 128 x 32"; PIPER_B200_ATT3=1 timeout -k 5 60 python tools/tap_errors.py medium 128 32; echo "rc=$?"
  nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv
} > gpurun_out/c9_att3.txt 2>&1
cut -c1-330 gpurun_out/c9_att3.txt
