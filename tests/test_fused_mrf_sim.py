"""The experimental fused MRF kernel's own body (piper_b200/csrc/mrf_fused_body.inl), compiled against a CPU model of the
hardware primitives (tests/sim/mrf_sim.cpp: every CTA thread a std::thread, blocking mbarriers, TMEM array, tcgen05.mma
decoded from its descriptors), against the oracle's generator stage.  No GPU involved: this checks index arithmetic,
operand layouts, the TMEM column map, TMA alignment and the barrier protocol - not the hardware's asynchrony."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle.voice_loader import load_voice
from oracle.vits_oracle import Oracle
from piper_b200 import _lib, voicegen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM = os.path.join(ROOT, "tests", "sim", "libmrf_sim.so")


def _plan_and_taps(path, stage, fuse_post=0):
    lib = _lib.load()
    plan = (C.c_int32 * 32)()
    wb, nb = C.c_int64(0), C.c_int64(0)
    _lib.check(lib.pb200_debug_mrf_pack(path.encode(), stage, fuse_post, plan, None, C.byref(wb), None, C.byref(nb)))
    w = np.zeros(wb.value + 64, np.uint8)
    w = w[(-w.ctypes.data) % 64:][:wb.value]                      # 64-byte aligned like the device buffer
    b = np.zeros(nb.value, np.float32)
    _lib.check(lib.pb200_debug_mrf_pack(path.encode(), stage, fuse_post, plan, w.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(wb),
                                        b.ctypes.data_as(C.POINTER(C.c_float)), C.byref(nb)))
    return plan, w, b


@pytest.mark.parametrize("arch,stage,n_phonemes,grid", [("tiny", 0, (12, 3, 7), 2), ("tiny-high", 0, (9, 2), 3),
                                                        ("tiny", 0, (1, 0, 12), 1),      # items shorter than one tile, one CTA
                                                        ("tiny", 0, (2, 12), 64)])       # more CTAs than tiles
def test_kernel_body_on_cpu_model_matches_oracle(lib_built, arch, stage, n_phonemes, grid):
    if not os.path.exists(SIM):
        import subprocess
        subprocess.run(["make", "-C", os.path.join(ROOT, "piper_b200", "csrc"), "../../tests/sim/libmrf_sim.so"], check=True,
                       stdout=subprocess.DEVNULL)
    sim = C.CDLL(SIM)
    path = voicegen.cached_voice(arch)
    plan, w, bias = _plan_and_taps(path, stage)
    assert plan[0] == 1
    spec, wd, attrs = load_voice(path)
    orc = Oracle(spec, wd, attrs)
    xs, refs = [], []
    for i, n in enumerate(n_phonemes):
        dump = {}
        orc.infer(voicegen.benchmark_ids(n, seed=20 + i), (0.667, 1.0, 0.8), dump=dump)
        xs.append(dump[f"up{stage}"].numpy())
        refs.append(dump[f"stage{stage}"].numpy())
    B = len(xs)
    lens = np.asarray([x.shape[1] for x in xs], np.int32)
    assert len(set(lens.tolist())) > 1 and lens.max() > plan[5]     # ragged, and more than one tile for the longest item
    cs = (int(lens.max()) + 3) & ~3
    rng = np.random.default_rng(0)
    x = rng.standard_normal((B, 32, cs)).astype(np.float32) * 50.0    # stale data past each item's length must not matter
    for b in range(B):
        x[b, :, :lens[b]] = xs[b]
    y = np.full((B, 32, cs), 7e7, np.float32)
    err = C.create_string_buffer(512)
    rc = sim.mrf_sim_run(x.ctypes.data_as(C.POINTER(C.c_float)), y.ctypes.data_as(C.POINTER(C.c_float)),
                         lens.ctypes.data_as(C.POINTER(C.c_int32)), B, C.c_longlong(32 * cs), cs, 1,
                         w.ctypes.data_as(C.POINTER(C.c_uint8)), bias.ctypes.data_as(C.POINTER(C.c_float)), plan,
                         int(lens.max()), grid, None, None, None, err, len(err))
    assert rc == 0, err.value.decode()
    for b in range(B):
        got, ref = y[b, :, :lens[b]], refs[b]
        e = float(np.abs(got - ref).max())
        assert e <= 1e-4 * max(1.0, float(np.abs(ref).max())), (b, e)
        assert np.all(y[b, :, lens[b]:] == 7e7), "the kernel stored outside the utterance"


def test_fused_generator_tail_matches_oracle_audio(lib_built):
    """Last stage of the medium generator with conv_post + tanh fused behind it: the kernel writes the waveform itself
    (tightly packed per utterance, like conv_post_kernel) and never the stage output."""
    if not os.path.exists(SIM):
        import subprocess
        subprocess.run(["make", "-C", os.path.join(ROOT, "piper_b200", "csrc"), "../../tests/sim/libmrf_sim.so"], check=True,
                       stdout=subprocess.DEVNULL)
    sim = C.CDLL(SIM)
    path = voicegen.cached_voice("medium")
    spec, wd, attrs = load_voice(path)
    stage = len(spec.up_rates) - 1
    plan0, _, _ = _plan_and_taps(path, stage, 0)
    plan, w, bias = _plan_and_taps(path, stage, 1)
    assert plan[0] == 1 and plan[27] == 7 and plan[4] == plan0[4] + 3 and plan[5] == plan0[5] - 6
    orc = Oracle(spec, wd, attrs)
    xs, audios = [], []
    for i, n in enumerate((1, 3)):
        dump = {}
        audio = orc.infer(voicegen.benchmark_ids(n, seed=40 + i), (0.667, 1.0, 0.8), dump=dump)
        xs.append(dump[f"up{stage}"].numpy())
        audios.append(np.asarray(audio, np.float32))
    B = len(xs)
    lens = np.asarray([x.shape[1] for x in xs], np.int32)
    assert all(len(a) == L for a, L in zip(audios, lens))
    cs = (int(lens.max()) + 3) & ~3
    x = np.random.default_rng(0).standard_normal((B, 32, cs)).astype(np.float32) * 50.0
    for b in range(B):
        x[b, :, :lens[b]] = xs[b]
    y = np.full((B, 32, cs), 7e7, np.float32)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    out = np.full(int(off[-1]) + 8, 9e9, np.float32)
    post = np.ascontiguousarray(wd["dec.conv_post.weight"][0], np.float32)          # [32][7]
    assert post.shape == (32, 7)
    err = C.create_string_buffer(512)
    rc = sim.mrf_sim_run(x.ctypes.data_as(C.POINTER(C.c_float)), y.ctypes.data_as(C.POINTER(C.c_float)),
                         lens.ctypes.data_as(C.POINTER(C.c_int32)), B, C.c_longlong(32 * cs), cs, 1,
                         w.ctypes.data_as(C.POINTER(C.c_uint8)), bias.ctypes.data_as(C.POINTER(C.c_float)), plan,
                         int(lens.max()), 5, post.ctypes.data_as(C.POINTER(C.c_float)), out.ctypes.data_as(C.POINTER(C.c_float)),
                         off.ctypes.data_as(C.POINTER(C.c_longlong)), err, len(err))
    assert rc == 0, err.value.decode()
    for b in range(B):
        got = out[off[b]:off[b + 1]]
        e = float(np.abs(got - audios[b]).max())
        assert e <= 1e-4, (b, e)                                             # waveform units (|audio| <= 1)
        assert float(np.sqrt((audios[b] ** 2).mean())) > 0.05
    assert np.all(out[off[-1]:] == 9e9) and np.all(y == 7e7), "wrote outside the packed audio / wrote the stage output"
