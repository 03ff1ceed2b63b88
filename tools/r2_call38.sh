#!/bin/bash
# round 2, call 38: converter / epilogue warp split per launch (PIPER_B200_V2_CW: 4 = previous build's fixed split, 8 = forced, unset = per-launch choice)
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
bench() { local name=$1; shift; env "$@" timeout 200 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/c38_$name.json 2> gpurun_out/c38_$name.err; echo "rc=$?"; python - <<PY
import json
d = json.load(open("gpurun_out/c38_$name.json")); print("$name", round(d["value"] / 1e6, 1), round(d["ms_per_step"], 3), {k: round(v, 3) for k, v in d.get("stage_ms", {}).items()}, "batch1", d.get("batch1", {}).get("latency_ms"))
PY
}
bench cw4 PIPER_B200_V2_CW=4
bench auto X=1
bench cw8 PIPER_B200_V2_CW=8
bench auto2 X=1
PIPER_B200_V2_CW=8 timeout 600 python -m pytest tests/test_gpu_conv_kernels.py -m gpu -x -q > gpurun_out/c38_kernels_cw8.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/c38_kernels_cw8.txt
timeout 600 python -m pytest tests/test_gpu_conv_kernels.py -m gpu -x -q > gpurun_out/c38_kernels.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/c38_kernels.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/c38_parity.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/c38_parity.txt
PIPER_B200_V2_CW=8 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/c38_parity_cw8.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/c38_parity_cw8.txt
PIPER_B200_PROF_ROLES=1 timeout 300 python tools/layer_report.py > gpurun_out/c38_roles.txt 2>&1; echo "rc=$?"
tail -8 gpurun_out/c38_roles.txt
