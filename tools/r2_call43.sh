#!/bin/bash
# round 2, call 43: the other BASELINE.json configurations on the final build (clock records inside each JSON line)
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 100 python bench.py --config 2 --no-cpu-baseline > gpurun_out/c43_cfg2.json 2> gpurun_out/c43_cfg2.err; echo "rc=$?"
timeout 150 python bench.py --config 4 --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/c43_cfg4.json 2> gpurun_out/c43_cfg4.err; echo "rc=$?"
timeout 100 python bench.py --config 5 --arch medium --no-cpu-baseline > gpurun_out/c43_cfg5_medium.json 2> gpurun_out/c43_cfg5_medium.err; echo "rc=$?"
timeout 150 python bench.py --config 5 --arch high --no-cpu-baseline > gpurun_out/c43_cfg5_high.json 2> gpurun_out/c43_cfg5_high.err; echo "rc=$?"
for f in cfg2 cfg4 cfg5_medium cfg5_high; do cut -c1-400 gpurun_out/c43_$f.json; echo; done
