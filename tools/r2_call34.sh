#!/bin/bash
# round 2, call 34 (new session): role-wait report and quick bench of HEAD on a fresh box
mkdir -p gpurun_out
timeout 200 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/c34_bench.json 2> gpurun_out/c34_bench.err; echo "rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/c34_bench.json")); print(d["value"], d["ms_per_step"], d.get("stage_ms"))
PY
PIPER_B200_PROF_ROLES=1 timeout 300 python tools/layer_report.py > gpurun_out/c34_roles.txt 2>&1; echo "rc=$?"
tail -5 gpurun_out/c34_roles.txt
