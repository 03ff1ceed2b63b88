#!/bin/bash
# round 2, call 40: A-stationary instantiations (default on) + short last query tiles of the attention on the CUDA cores
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
bench() { local name=$1; shift; env "$@" timeout 200 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/c40_$name.json 2> gpurun_out/c40_$name.err; echo "rc=$?"; python - <<PY
import json
d = json.load(open("gpurun_out/c40_$name.json")); print("$name", round(d["value"] / 1e6, 1), round(d["ms_per_step"], 3), {k: round(v, 3) for k, v in d.get("stage_ms", {}).items()}, "batch1", d.get("batch1", {}).get("latency_ms"))
PY
}
bench default X=1
bench notail PIPER_B200_ATT_TAIL=0
bench noastat PIPER_B200_V2_ASTAT=0
bench neither PIPER_B200_V2_ASTAT=0 PIPER_B200_ATT_TAIL=0
bench default2 X=1
timeout 600 python -m pytest tests/test_gpu_conv_kernels.py -m gpu -x -q > gpurun_out/c40_kernels.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/c40_kernels.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/c40_parity.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/c40_parity.txt
PIPER_B200_PROF_ROLES=1 timeout 300 python tools/layer_report.py > gpurun_out/c40_roles.txt 2>&1; echo "rc=$?"
tail -8 gpurun_out/c40_roles.txt
