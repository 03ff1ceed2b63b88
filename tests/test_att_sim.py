"""The experimental tensor-core relative-position attention (piper_b200/csrc/att_body.inl) on the CPU model of the
primitives (tests/sim) against a direct statement of attentions.py:225-272 in torch.  No GPU involved."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM = os.path.join(ROOT, "tests", "sim", "libatt_sim.so")


@pytest.fixture(scope="module")
def sim():
    if not os.path.exists(SIM):
        subprocess.run(["make", "-C", os.path.join(ROOT, "piper_b200", "csrc"), "../../tests/sim/libatt_sim.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return C.CDLL(SIM)


def _reference(q, k, v, rel_k, rel_v, window=4):
    """q, k, v [dk][T] of one head; emb_rel_k / emb_rel_v [2w+1][dk] shared across heads -> [dk][T]."""
    dk, T = q.shape
    qs = (q / np.sqrt(np.float32(dk))).double().t()                     # query / sqrt(k_channels) first (attentions.py:232)
    S = qs @ k.double()
    i = torch.arange(T)[:, None]
    j = torch.arange(T)[None, :]
    rel = j - i + window
    band = (rel >= 0) & (rel <= 2 * window)
    logits = qs @ rel_k.double().t()                                     # [T][2w+1]
    S = S + torch.where(band, logits.gather(1, rel.clamp(0, 2 * window)), torch.zeros((), dtype=torch.float64))
    P = torch.softmax(S, dim=1)
    O = P @ v.double().t()                                               # [T][dk]
    Pb = torch.where(band, P, torch.zeros((), dtype=torch.float64))
    for r in range(2 * window + 1):
        jj = torch.arange(T) + r - window
        ok = (jj >= 0) & (jj < T)
        O[ok] += Pb[torch.arange(T)[ok], jj[ok]][:, None] * rel_v[r].double()[None, :]
    return O.t().float()


@pytest.mark.parametrize("mode", ["rows", "tm", "tm_flat"])     # per-row bulk copies / tensor-map boxes / tensor map over the flat layout
@pytest.mark.parametrize("H,n_heads,lens", [(192, 2, (259, 17)),       # medium: dk = 96, five key blocks, ragged
                                             (96, 2, (130,)),           # x-low: dk = 48
                                             (32, 2, (64, 1, 65)),      # dk = 16; exactly one block, one key, one past a block
                                             (64, 4, (600,))])          # four heads, five query tiles, ten key blocks
def test_tensor_core_attention_on_cpu_model(sim, H, n_heads, lens, mode):
    B, dk = len(lens), H // n_heads
    Tmax = max(lens)
    cs = ((Tmax + 3) & ~3) + (4 if mode == "tm_flat" else 0)
    rng = np.random.default_rng(H + Tmax)
    qkv = rng.standard_normal((B, 3 * H, cs)).astype(np.float32) * 30.0   # stale data past each item's length
    for b, T in enumerate(lens):
        qkv[b, :, :T] = rng.standard_normal((3 * H, T)).astype(np.float32) * np.float32(1.5)
    rel_k = (rng.standard_normal((9, dk)) * 0.3).astype(np.float32)
    rel_v = (rng.standard_normal((9, dk)) * 0.3).astype(np.float32)
    out = np.full((B, H, cs), 7e7, np.float32)
    lens_a = np.asarray(lens, np.int32)
    err = C.create_string_buffer(512)
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    flat = mode == "tm_flat"
    qkv_dev = np.ascontiguousarray(qkv.transpose(1, 0, 2)) if flat else qkv          # [channel][item][slot]
    out_dev = np.ascontiguousarray(out.transpose(1, 0, 2)) if flat else out
    rc = sim.att_sim_run(fp(qkv_dev), fp(out_dev), fp(rel_k), fp(rel_v), lens_a.ctypes.data_as(C.POINTER(C.c_int32)), B, H, n_heads, cs,
                         Tmax, err, len(err), 0 if mode == "rows" else 1, 1 if flat else 0)
    assert rc == 0, err.value.decode()
    if flat:
        out = out_dev.transpose(1, 0, 2)
    for b, T in enumerate(lens):
        for h in range(n_heads):
            q = torch.from_numpy(qkv[b, h * dk:(h + 1) * dk, :T])
            k = torch.from_numpy(qkv[b, H + h * dk:H + (h + 1) * dk, :T])
            v = torch.from_numpy(qkv[b, 2 * H + h * dk:2 * H + (h + 1) * dk, :T])
            ref = _reference(q, k, v, torch.from_numpy(rel_k), torch.from_numpy(rel_v))
            got = torch.from_numpy(np.ascontiguousarray(out[b, h * dk:(h + 1) * dk, :T]))
            e = float((got - ref).abs().max())
            assert e <= 2e-5 * max(1.0, float(ref.abs().max())), (b, h, e)
        assert np.all(out[b, :, T:] == 7e7), "stored outside the utterance"


def test_short_last_tiles_are_left_to_the_tail_kernel(sim):
    """Tail mode of the launcher (att_mma.cu: more query tiles than SMs): a LAST tile of at most tail_thr rows is skipped -
    its rows stay untouched for encoder.cu's rel_attention_tail_kernel - every other tile is computed as before."""
    H, n_heads, lens, thr = 32, 2, (259, 131, 200, 16, 144), 16
    B, dk = len(lens), H // n_heads
    Tmax = max(lens)
    cs = (Tmax + 3) & ~3
    rng = np.random.default_rng(5)
    qkv = rng.standard_normal((B, 3 * H, cs)).astype(np.float32) * 30.0
    for b, T in enumerate(lens):
        qkv[b, :, :T] = rng.standard_normal((3 * H, T)).astype(np.float32) * np.float32(1.5)
    rel_k = (rng.standard_normal((9, dk)) * 0.3).astype(np.float32)
    rel_v = (rng.standard_normal((9, dk)) * 0.3).astype(np.float32)
    out = np.full((B, H, cs), 7e7, np.float32)
    lens_a = np.asarray(lens, np.int32)
    err = C.create_string_buffer(512)
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    rc = sim.att_sim_run2(fp(qkv), fp(out), fp(rel_k), fp(rel_v), lens_a.ctypes.data_as(C.POINTER(C.c_int32)), B, H, n_heads, cs, Tmax,
                          err, len(err), 1, 0, thr)
    assert rc == 0, err.value.decode()
    for b, T in enumerate(lens):
        q0 = (T - 1) // 128 * 128
        skipped = q0 > 0 and T - q0 <= thr                               # 259 -> rows 256.., 131 -> rows 128.., 144 -> rows 128..; 200 and 16 are computed
        done = q0 if skipped else T
        assert np.all(out[b, :, done:] == 7e7), (b, "rows of a skipped tile / past the utterance were written")
        for h in range(n_heads):
            q = torch.from_numpy(qkv[b, h * dk:(h + 1) * dk, :T])
            k = torch.from_numpy(qkv[b, H + h * dk:H + (h + 1) * dk, :T])
            v = torch.from_numpy(qkv[b, 2 * H + h * dk:2 * H + (h + 1) * dk, :T])
            ref = _reference(q, k, v, torch.from_numpy(rel_k), torch.from_numpy(rel_v))
            got = torch.from_numpy(np.ascontiguousarray(out[b, h * dk:(h + 1) * dk, :done]))
            assert float((got - ref[:, :done]).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max())), (b, h)
    assert any((T - 1) // 128 * 128 > 0 and T - (T - 1) // 128 * 128 <= thr for T in lens)
