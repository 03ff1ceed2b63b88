#!/bin/bash
# round 2, call 37: conv2 epilogue second revision (no register copies, cross-tile prefetch, paired tcgen05.ld, one base
# address per tensor and item) against HEAD's library (_ab/libpiper_b200_head.so), kernel-level GPU tests, role report
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
bench() { timeout 200 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/c37_$1.json 2> gpurun_out/c37_$1.err; echo "rc=$?"; python - <<PY
import json
d = json.load(open("gpurun_out/c37_$1.json")); print("$1", round(d["value"] / 1e6, 1), round(d["ms_per_step"], 3), {k: round(v, 3) for k, v in d.get("stage_ms", {}).items()})
PY
}
cp piper_b200/libpiper_b200.so _ab/libpiper_b200_new.so
bench new1
cp _ab/libpiper_b200_head.so piper_b200/libpiper_b200.so
bench head
cp _ab/libpiper_b200_new.so piper_b200/libpiper_b200.so
bench new2
timeout 600 python -m pytest tests/test_gpu_conv_kernels.py -m gpu -x -q > gpurun_out/c37_kernels.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/c37_kernels.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/c37_parity.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/c37_parity.txt
PIPER_B200_PROF_ROLES=1 timeout 300 python tools/layer_report.py > gpurun_out/c37_roles.txt 2>&1; echo "rc=$?"
tail -8 gpurun_out/c37_roles.txt
