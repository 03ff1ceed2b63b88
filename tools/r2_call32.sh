#!/bin/bash
# round 2, call 32: tcgen05.mma pacing probe with a uniform-register issue loop
mkdir -p gpurun_out
timeout 120 tools/probe/mma_probe 64 0 7 > gpurun_out/c32_probe64.txt 2>&1; echo "rc=$?"
timeout 120 tools/probe/mma_probe 128 0 7 > gpurun_out/c32_probe128.txt 2>&1; echo "rc=$?"
timeout 120 tools/probe/mma_probe 32 0 7 > gpurun_out/c32_probe32.txt 2>&1; echo "rc=$?"
timeout 60 tools/probe/mma_probe 64 8 8 > gpurun_out/c32_probe64_sw.txt 2>&1; echo "rc=$?"
cat gpurun_out/c32_probe64.txt | head -48
