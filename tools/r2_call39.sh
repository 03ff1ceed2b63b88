#!/bin/bash
# round 2, call 39: A-stationary order with the whole converted window resident (PIPER_B200_V2_ASTAT=1), converter split by
# the host's cost rule (PIPER_B200_V2_CS=1)
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
bench() { local name=$1; shift; env "$@" timeout 200 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/c39_$name.json 2> gpurun_out/c39_$name.err; echo "rc=$?"; python - <<PY
import json
d = json.load(open("gpurun_out/c39_$name.json")); print("$name", round(d["value"] / 1e6, 1), round(d["ms_per_step"], 3), {k: round(v, 3) for k, v in d.get("stage_ms", {}).items()}, "batch1", d.get("batch1", {}).get("latency_ms"))
PY
}
bench base X=1
bench astat PIPER_B200_V2_ASTAT=1
bench cs PIPER_B200_V2_CS=1
bench both PIPER_B200_V2_ASTAT=1 PIPER_B200_V2_CS=1
bench base2 X=1
PIPER_B200_V2_ASTAT=1 PIPER_B200_V2_CS=1 timeout 600 python -m pytest tests/test_gpu_conv_kernels.py -m gpu -x -q > gpurun_out/c39_kernels_astat.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/c39_kernels_astat.txt
PIPER_B200_V2_ASTAT=1 PIPER_B200_V2_CS=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/c39_parity_astat.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/c39_parity_astat.txt
PIPER_B200_V2_ASTAT=1 PIPER_B200_PROF_ROLES=1 timeout 300 python tools/layer_report.py > gpurun_out/c39_roles_astat.txt 2>&1; echo "rc=$?"
tail -8 gpurun_out/c39_roles_astat.txt
