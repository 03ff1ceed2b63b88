#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
t() { local name=$1; shift; echo "== $name: $*" >> gpurun_out/d2_taps.txt; env "$@" timeout 200 python tools/tap_errors.py tiny 20 >> gpurun_out/d2_taps.txt 2>&1; env "$@" timeout 200 python tools/tap_errors.py medium 64 >> gpurun_out/d2_taps.txt 2>&1; env "$@" timeout 300 python tools/tap_errors.py medium 128 32 >> gpurun_out/d2_taps.txt 2>&1; }
t base X=1
t v2_1 PIPER_B200_V2=1
t v2_2 PIPER_B200_V2=2
t v2_2_dec PIPER_B200_V2=2 PIPER_B200_MMA=1
t v2_2_flow PIPER_B200_V2=2 PIPER_B200_MMA=2
t v2_2_enc PIPER_B200_V2=2 PIPER_B200_MMA=4
t v2_2_dp PIPER_B200_V2=2 PIPER_B200_MMA=8
cat gpurun_out/d2_taps.txt | grep -v "^==" | cut -c1-400 | tail -40
echo "== sanitizer on the tensor-map path" > gpurun_out/d2_sanitizer.txt
cat > /tmp/one_tm.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np
from piper_b200 import engine
rng = np.random.default_rng(0)
x = rng.standard_normal((2, 32, 300)).astype(np.float32); w = (rng.standard_normal((32, 32, 3)) / 10).astype(np.float32)
y = engine.debug_conv1d(3, x, w, None, 1, 0.0, None)
print("ok", float(np.abs(y).max()))
PY
PIPER_B200_V2=2 PIPER_B200_V2_TM=1 timeout 300 compute-sanitizer --tool memcheck python /tmp/one_tm.py >> gpurun_out/d2_sanitizer.txt 2>&1
tail -40 gpurun_out/d2_sanitizer.txt | cut -c1-300
