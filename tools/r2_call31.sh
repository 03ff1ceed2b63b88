#!/bin/bash
# round 2, call 31: tcgen05.mma pacing probe + stacked vs three-instruction form in the real step
mkdir -p gpurun_out
timeout 120 tools/probe/mma_probe 64 0 7 > gpurun_out/c31_probe64.txt 2>&1; echo "rc=$?"
timeout 120 tools/probe/mma_probe 128 0 7 > gpurun_out/c31_probe128.txt 2>&1; echo "rc=$?"
timeout 60 tools/probe/mma_probe 64 8 8 > gpurun_out/c31_probe64_sw.txt 2>&1; echo "rc=$?"
nvidia-smi --query-gpu=name,clocks.sm --format=csv,noheader
PIPER_B200_V2_MMA3=1 timeout 200 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/c31_bench_mma3.json 2> gpurun_out/c31_bench_mma3.err; echo "rc=$?"
timeout 200 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/c31_bench.json 2> gpurun_out/c31_bench.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("c31_bench_mma3", "c31_bench"):
    d = json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"], d["ms_per_step"])
PY
head -40 gpurun_out/c31_probe64.txt
