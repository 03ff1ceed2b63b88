"""Tile-level prototype (CPU, torch) of the fused MRF stage planned for the generator.

One output tile [t0, t0 + TO) of  x -> (sum_j ResBlock_j(x)) / n  (models.py:356-363) is computed from a single staged
window of x, the way the fused CUDA kernel will: every convolution of a resblock chain runs on a window that shrinks by
its own half-width, intermediates never leave the tile.  The one thing a fused kernel gets wrong if it is careless is
the padding: the reference zero-pads the *input of every conv* at the utterance edges, so an intermediate tensor must
read as 0 outside [0, L) even though the halo recompute would happily produce bias + conv(inside values) there.
`mask=False` reproduces that bug; tests/test_fused_mrf_proto.py pins both behaviours against the oracle.

A chain is a list of steps (weight [co, ci, k], bias, dilation, residual_from): residual_from = index of the earlier
chain tensor added after this conv (0 = the chain input), or None (ResBlock1's first conv of a pair).
"""
from typing import List, Optional, Sequence, Tuple
import torch
import torch.nn.functional as F

Step = Tuple[torch.Tensor, Optional[torch.Tensor], int, Optional[int]]


def resblock_chains(w: dict, spec, stage: int) -> List[List[Step]]:
    """Chains of the three (n) resblocks that follow upsample stage `stage`, from an oracle weight dict."""
    nk = len(spec.rb_kernels)
    out = []
    for j in range(nk):
        rb = f"dec.resblocks.{stage * nk + j}"
        steps: List[Step] = []
        for c, d in enumerate(spec.rb_dilations[j]):
            if spec.resblock == 1:       # y = y + conv2(lrelu(conv1(lrelu(y))))   (modules.py ResBlock1)
                src = len(steps)         # index of y in the chain's tensor list
                steps.append((w[f"{rb}.convs1.{c}.weight"], w.get(f"{rb}.convs1.{c}.bias"), d, None))
                steps.append((w[f"{rb}.convs2.{c}.weight"], w.get(f"{rb}.convs2.{c}.bias"), 1, src))
            else:                        # y = y + conv(lrelu(y))                  (modules.py ResBlock2)
                steps.append((w[f"{rb}.convs.{c}.weight"], w.get(f"{rb}.convs.{c}.bias"), d, len(steps)))
        out.append(steps)
    return out


def halo(chain: Sequence[Step]) -> int:
    return sum((s[0].shape[2] - 1) // 2 * s[2] for s in chain)


def fused_tile(x: torch.Tensor, chains: Sequence[Sequence[Step]], t0: int, TO: int, mask: bool = True) -> torch.Tensor:
    """x [C, L] (one utterance).  Returns y[:, t0 : min(t0 + TO, L)]."""
    C, L = x.shape
    H = max(halo(c) for c in chains)
    lo, hi = t0 - H, t0 + TO + H
    win = torch.zeros(C, hi - lo)                                   # staged window: zeros outside the utterance
    a, b = max(lo, 0), min(hi, L)
    win[:, a - lo:b - lo] = x[:, a:b]
    total = None
    for chain in chains:
        # tensors[i] covers positions [start[i], start[i] + tensors[i].shape[1])
        tensors, start = [win], [lo]
        h_left = halo(chain)
        # the chain input is needed on [t0 - h_left, t0 + TO + h_left): crop the shared window to it
        cur, cur_start = win[:, H - h_left:win.shape[1] - (H - h_left)], t0 - h_left
        tensors[0], start[0] = cur, cur_start
        for (W, bias, dil, res) in chain:
            k = W.shape[2]
            hw = (k - 1) // 2 * dil
            y = F.conv1d(F.leaky_relu(cur, 0.1)[None], W, bias, dilation=dil)[0]      # 'valid': shrinks by hw each side
            y_start = cur_start + hw
            if res is not None:
                r, rs = tensors[res], start[res]
                y = y + r[:, y_start - rs:y_start - rs + y.shape[1]]
            if mask:                                                  # an intermediate is 0 outside the utterance
                pos = torch.arange(y_start, y_start + y.shape[1])
                y = y * ((pos >= 0) & (pos < L)).to(y.dtype)[None]
            tensors.append(y)
            start.append(y_start)
            cur, cur_start = y, y_start
        assert cur_start == t0 and cur.shape[1] == TO
        total = cur if total is None else total + cur
    out = total / len(chains)
    return out[:, :max(0, min(TO, L - t0))]


def fused_stage(x: torch.Tensor, chains, TO: int = 184, mask: bool = True) -> torch.Tensor:
    C, L = x.shape
    return torch.cat([fused_tile(x, chains, t0, TO, mask) for t0 in range(0, L, TO)], 1)
