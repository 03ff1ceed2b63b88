"""Kernel-level check of the second-generation conv (conv_mma2.cu) on the GPU against float64 torch: every case x
precision, errors printed per case (one failure does not stop the rest).  PIPER_B200_V2=2 lets small launches through;
PIPER_B200_V2_TM=1 selects the tensor-map loads.   usage: PIPER_B200_V2=2 python tools/conv2_check.py"""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
from piper_b200 import engine

def ref(x, w, b, dil, slope, resid):
    xt = torch.from_numpy(x).double()
    if slope:
        xt = F.leaky_relu(xt, slope)
    y = F.conv1d(xt, torch.from_numpy(w).double(), None if b is None else torch.from_numpy(b).double(), dilation=dil,
                 padding=dil * (w.shape[2] - 1) // 2)
    if resid is not None:
        y = y + torch.from_numpy(resid).double()
    return y.numpy()

CASES = [(2, 32, 32, 3, 1, 300), (1, 32, 32, 7, 12, 1000), (2, 64, 64, 5, 6, 517), (1, 128, 128, 7, 3, 260), (1, 256, 256, 3, 1, 200),
         (2, 192, 384, 5, 1, 519), (1, 192, 768, 3, 1, 259), (1, 768, 192, 3, 1, 259), (2, 96, 192, 1, 1, 600), (1, 192, 96, 1, 1, 333),
         (1, 192, 576, 1, 1, 259), (1, 192, 192, 1, 1, 259), (32, 192, 192, 1, 1, 259), (32, 192, 576, 1, 1, 259),
         (4, 32, 32, 7, 12, 40000), (2, 64, 64, 5, 6, 45001), (2, 128, 128, 7, 3, 20003), (2, 192, 384, 5, 1, 6001), (1, 192, 768, 3, 1, 5000)]
bad = 0
for case in CASES:
    B, ci, co, k, dil, L = case
    rng = np.random.default_rng(hash(case) % (2 ** 31))
    x = rng.standard_normal((B, ci, L)).astype(np.float32) * 2
    w = (rng.standard_normal((co, ci, k)) / np.sqrt(ci * k)).astype(np.float32)
    b = rng.standard_normal(co).astype(np.float32)
    r = rng.standard_normal((B, co, L)).astype(np.float32)
    want = ref(x, w, b, dil, 0.1, r)
    for be, name in ((3, "bf16x3"), (4, "tf32x3/2"), (5, "f16x3/1"), (6, "f16x3/2")):
        try:
            y = engine.debug_conv1d(be, x, w, b, dil, 0.1, r)
            e = np.abs(y - want)
            tol = 3e-4
            flag = "ok " if e.max() <= tol else "BAD"
            bad += flag == "BAD"
            msg = f"{flag} {name:9s} {case} max err {e.max():.3e}"
            if flag == "BAD":
                idx = np.argwhere(e > tol)
                ts = np.unique(idx[:, 2])
                msg += f" n_bad {len(idx)} t range {ts.min()}..{ts.max()} ({len(ts)} cols) rows {np.unique(idx[:, 1])[:8].tolist()} items {np.unique(idx[:, 0]).tolist()} t%128 {np.unique(ts % 128)[:12].tolist()}"
            print(msg, flush=True)
        except Exception as ex:
            bad += 1
            print(f"EXC {name:9s} {case} {str(ex)[:200]}", flush=True)
print("total bad", bad)
