#!/bin/bash
# round 2, call 41: key-parallel CUDA-core kernel for the short last query tiles of the attention
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
bench() { local name=$1; shift; env "$@" timeout 200 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/c41_$name.json 2> gpurun_out/c41_$name.err; echo "rc=$?"; python - <<PY
import json
d = json.load(open("gpurun_out/c41_$name.json")); print("$name", round(d["value"] / 1e6, 1), round(d["ms_per_step"], 3), {k: round(v, 3) for k, v in d.get("stage_ms", {}).items()}, "batch1", d.get("batch1", {}).get("latency_ms"))
PY
}
bench default X=1
bench notail PIPER_B200_ATT_TAIL=0
bench default2 X=1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/c41_parity.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/c41_parity.txt
timeout 300 python - > gpurun_out/c41_tail_check.txt 2>&1 <<'PY'
# tails of 1 .. 16 rows behind one or two full tiles, 40 utterances (more query tiles than SMs): tensor-core + tail kernels
# against the all-tensor-core path of a second process is not possible in one process (switch read once), so compare with the oracle
import numpy as np, sys
sys.path.insert(0, ".")
from piper_b200 import engine, voicegen
from oracle.voice_loader import load_voice
from oracle.vits_oracle import Oracle
path = voicegen.cached_voice("medium")
v = engine.Voice(path, 0)
spec, w, attrs = load_voice(path)
orc = Oracle(spec, w, attrs)
rng = np.random.default_rng(3)
n_ph = [64 + (i % 9) for i in range(40)]           # ids = 2 n + 3 -> 131 .. 147 ids: one full tile + 3 .. 19 rows
ids = [voicegen.benchmark_ids(n, seed=50 + i) for i, n in enumerate(n_ph)]
eps_dp = [rng.standard_normal((2, len(i))).astype(np.float32) for i in ids]
Tz = 6 * max(len(i) for i in ids)
eps_z = rng.standard_normal((len(ids), 192, Tz)).astype(np.float32)
wavs, sec = v.synthesize_batch(ids, (0.667, 1.0, 0.8), eps_dp=eps_dp, eps_z=eps_z)
worst = 0.0
for b in (0, 1, 5, 8, 17, 39):
    r = orc.infer(ids[b], (0.667, 1.0, 0.8), eps_dp[b], eps_z[b])
    assert wavs[b].shape == r.shape, (b, wavs[b].shape, r.shape)
    e = float(np.abs(wavs[b] - r).max()); worst = max(worst, e)
    print(b, len(ids[b]), e)
print("worst", worst)
assert worst <= 1e-3
PY
echo "rc=$?"; tail -8 gpurun_out/c41_tail_check.txt
