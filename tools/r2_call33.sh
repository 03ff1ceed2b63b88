#!/bin/bash
# round 2, call 33: integer divisions removed from the per-tap / per-row loops of the conv kernel
mkdir -p gpurun_out
timeout 200 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/c33_bench.json 2> gpurun_out/c33_bench.err; echo "rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/c33_bench.json")); print(d["value"], d["ms_per_step"], d.get("stage_ms"))
PY
PIPER_B200_PROF_ROLES=1 timeout 300 python tools/layer_report.py > gpurun_out/c33_roles.txt 2>&1; echo "rc=$?"
tail -8 gpurun_out/c33_roles.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c33_pytest.txt 2>&1; echo "rc=$?"
tail -3 gpurun_out/c33_pytest.txt
