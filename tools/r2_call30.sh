#!/bin/bash
# round 2, call 30: per-role wait counters of every conv launch (config 3 step)
mkdir -p gpurun_out
PIPER_B200_PROF_ROLES=1 timeout 300 python tools/layer_report.py > gpurun_out/c30_roles.txt 2>&1; echo "rc=$?"
cp gpurun_out/layer_report.json gpurun_out/c30_roles.json
timeout 200 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/c30_bench.json 2> gpurun_out/c30_bench.err; echo "rc=$?"
tail -c 600 gpurun_out/c30_bench.json
grep -c roles gpurun_out/c30_roles.txt
