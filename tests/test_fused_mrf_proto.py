"""The tile-level algorithm of the planned fused MRF stage (tools/fused_mrf_proto.py) against the oracle's generator:
halo recompute + zero-masked intermediates reproduce the layer-wise result; without the mask they do not."""
import os, sys
import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from oracle.voice_loader import load_voice
from oracle.vits_oracle import Oracle
from piper_b200 import voicegen
import fused_mrf_proto as fp


@pytest.mark.parametrize("arch", ["tiny", "tiny-high"])       # ResBlock2 and ResBlock1 generators
@pytest.mark.parametrize("TO", [184, 64, 1000])
def test_fused_tiles_match_layerwise(arch, TO):
    spec, w, attrs = load_voice(voicegen.cached_voice(arch))
    o = Oracle(spec, w, attrs)
    ids = voicegen.benchmark_ids(6, seed=3)
    dump = {}
    o.infer(ids, (0.667, 1.0, 0.8), dump=dump)
    tw = {k: torch.as_tensor(v) for k, v in w.items() if k.startswith("dec.resblocks")}
    for stage in range(len(spec.up_rates)):
        x, ref = dump[f"up{stage}"], dump[f"stage{stage}"]
        chains = fp.resblock_chains(tw, spec, stage)
        got = fp.fused_stage(x, chains, TO=TO)
        assert got.shape == ref.shape
        scale = float(ref.abs().max())
        assert float((got - ref).abs().max()) <= 2e-5 * max(1.0, scale), (arch, stage)
        # the padding pitfall: unmasked intermediates differ at the utterance edges (and only there)
        bad = fp.fused_stage(x, chains, TO=TO, mask=False)
        err = (bad - ref).abs().amax(0)
        H = max(fp.halo(c) for c in chains)
        assert float(err.max()) > 1e-4 * max(1.0, scale)
        if x.shape[1] > 2 * H + 2:
            assert float(err[H:x.shape[1] - H].max()) <= 2e-5 * max(1.0, scale)


def test_halos_of_the_piper_presets():
    class S: pass
    mk = lambda k: torch.zeros(4, 4, k)
    # medium / x-low / low: ResBlock2, kernels (3,5,7), dilations (1,2),(2,6),(3,12)  
    chains = [[(mk(3), None, 1, 0), (mk(3), None, 2, 1)], [(mk(5), None, 2, 0), (mk(5), None, 6, 1)],
              [(mk(7), None, 3, 0), (mk(7), None, 12, 1)]]
    assert [fp.halo(c) for c in chains] == [3, 16, 45]
    # high: ResBlock1, kernels (3,7,11), c1 dilations (1,3,5), c2 dilation 1
    hi = []
    for k in (3, 7, 11):
        c = []
        for d in (1, 3, 5):
            c += [(mk(k), None, d, None), (mk(k), None, 1, len(c))]
        hi.append(c)
    assert [fp.halo(c) for c in hi] == [12, 36, 60]


# ---- the packed operands of the experimental CUDA kernel (csrc/mrf_fused.cu), emulated on the CPU --------------------
def _bf16(t):                 # operand rounding of the kernel: FP16 (fp16x3 split)
    return t.to(torch.float16).to(torch.float32)


def _emulate_kernel_stage(x, plan, w_bytes, bias):
    """Functional model of mrf_fused_kernel: 256-row tiles with one row <-> position mapping for every conv, guard rows,
    zero-masked operands, stacked [W_hi ; W_lo] tap tiles consumed step-major, fp16x3 products (exact accumulation)."""
    ok, n_chains, n_steps, pair, hv, TO = plan[:6]
    k = plan[6:9]
    dil = np.asarray(plan[9:27]).reshape(3, 6)
    G0, GA, M, C = 12, 36, 256, 32
    taps = np.frombuffer(w_bytes, np.uint16).reshape(-1, 4, 64, 8)              # [tap][ci / 8][hi co | lo co][ci % 8]
    taps = torch.from_numpy(taps.view(np.float16).astype(np.float32))
    taps = taps.permute(0, 2, 1, 3).reshape(-1, 64, 32)                            # [tap][row][ci]
    bias = torch.from_numpy(bias.reshape(n_steps, n_chains, C).copy())
    L = x.shape[1]
    y = torch.zeros_like(x)
    for t0 in range(0, L, TO):
        p0 = t0 - hv
        pos0 = torch.arange(p0 - G0, p0 + M + G0)
        live0 = (pos0 >= 0) & (pos0 < L)
        win = torch.zeros(C, M + 2 * G0)
        win[:, live0] = x[:, pos0[live0]]
        a0 = torch.nn.functional.leaky_relu(win, 0.1)
        A0 = (_bf16(a0), _bf16(a0 - _bf16(a0)))
        AC = [[torch.zeros(C, M + 2 * GA), torch.zeros(C, M + 2 * GA)] for _ in range(n_chains)]
        carrier = [None] * n_chains
        pos = torch.arange(p0, p0 + M)
        inside = ((pos >= 0) & (pos < L)).float()
        total = torch.zeros(C, M)
        tap = 0
        for s in range(n_steps):
            closes = s % pair == pair - 1
            for c in range(n_chains):
                d = int(dil[c, s])
                hw = (k[c] - 1) // 2 * d
                (Ah, Al), guard = (A0, G0) if s == 0 else (AC[c], GA)
                acc = torch.zeros(M, C, dtype=torch.float64)
                for j in range(k[c]):
                    Wh, Wl = taps[tap, :32].double(), taps[tap, 32:].double()     # [co][ci]
                    tap += 1
                    r0 = guard - hw + j * d
                    ah, al = Ah[:, r0:r0 + M].t().double(), Al[:, r0:r0 + M].t().double()
                    acc += ah @ Wh.t() + (ah @ Wl.t() + al @ Wh.t())
                v = acc.float().t() + bias[s, c][:, None]
                if closes:
                    v = v + (win[:, G0:G0 + M] if s == pair - 1 else carrier[c])
                if s == n_steps - 1:
                    total = total + v
                else:
                    if closes:
                        carrier[c] = v
                    op = torch.nn.functional.leaky_relu(v, 0.1) * inside[None]
                    AC[c][0][:, GA:GA + M] = _bf16(op)
                    AC[c][1][:, GA:GA + M] = _bf16(op - _bf16(op))
        n = min(TO, L - t0)
        y[:, t0:t0 + n] = (total / n_chains)[:, hv:hv + n]
        assert tap == taps.shape[0]
    return y


@pytest.mark.parametrize("arch,stage", [("tiny", 0), ("tiny-high", 0)])
def test_fused_kernel_plan_and_packed_taps(lib_built, arch, stage):
    import ctypes as C
    from piper_b200 import _lib
    lib = _lib.load()
    path = voicegen.cached_voice(arch)
    plan = (C.c_int32 * 32)()
    wb, nb = C.c_int64(0), C.c_int64(0)
    _lib.check(lib.pb200_debug_mrf_pack(path.encode(), stage, 0, plan, None, C.byref(wb), None, C.byref(nb)))
    plan_l = list(plan)
    assert plan_l[0] == 1, "a 32-channel stage of a piper preset must be plannable"
    spec, w, attrs = load_voice(path)
    k = spec.rb_kernels
    pair = 2 if spec.resblock == 1 else 1
    assert plan_l[1:4] == [len(k), pair * len(spec.rb_dilations[0]), pair] and plan_l[6:9] == list(k)
    # hv = summed half-widths of every conv but the first, worst chain; TO = 256 - 2 hv
    hv = 0
    for j, kk in enumerate(k):
        ds = []
        for d in spec.rb_dilations[j]:
            ds += [d] if pair == 1 else [d, 1]
        hv = max(hv, sum((kk - 1) // 2 * d for d in ds[1:]))
    assert plan_l[4] == hv and plan_l[5] == 256 - 2 * hv
    assert wb.value == 4096 * sum(k) * plan_l[2] and nb.value == plan_l[2] * len(k) * 32
    wbuf = np.zeros(wb.value, np.uint8)
    bbuf = np.zeros(nb.value, np.float32)
    _lib.check(lib.pb200_debug_mrf_pack(path.encode(), stage, 0, plan, wbuf.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(wb),
                                        bbuf.ctypes.data_as(C.POINTER(C.c_float)), C.byref(nb)))
    o = Oracle(spec, w, attrs)
    dump = {}
    o.infer(voicegen.benchmark_ids(12, seed=5), (0.667, 1.0, 0.8), dump=dump)
    x, ref = dump[f"up{stage}"], dump[f"stage{stage}"]
    assert x.shape[0] == 32 and x.shape[1] > plan_l[5]                 # more than one tile, ragged last tile
    got = _emulate_kernel_stage(x, plan_l, wbuf.tobytes(), bbuf)
    err = float((got - ref).abs().max())
    assert err <= 1e-4 * max(1.0, float(ref.abs().max())), err


def test_fused_kernel_declines_other_widths(lib_built):
    import ctypes as C
    from piper_b200 import _lib
    lib = _lib.load()
    plan = (C.c_int32 * 32)()
    wb, nb = C.c_int64(0), C.c_int64(0)
    _lib.check(lib.pb200_debug_mrf_pack(voicegen.cached_voice("tiny").encode(), 1, 0, plan, None, C.byref(wb), None, C.byref(nb)))
    assert plan[0] == 0 and wb.value == 0                                # 16 channels: stays on the layer-wise path
