"""CPU tests that pin the oracle: against the reference's own PyTorch source (build container) and
against the golden fixtures minted from it (everywhere)."""
import glob
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, real_fixture_lines, real_voice_path
from piper_b200 import voicegen


def _oracle_for(tag):
    from oracle.voice_loader import load_voice
    from oracle.vits_oracle import Oracle
    if tag.startswith("real:"):
        path = real_voice_path()
        if path is None:
            pytest.skip("reference test voice not staged")
    else:
        _, arch, seed = tag.split(":")
        path = voicegen.cached_voice(arch, int(seed))
    spec, w, attrs = load_voice(path)
    return Oracle(spec, w, attrs), w


@pytest.mark.parametrize("name", sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
                                        if not os.path.basename(p).startswith(("stream_", "int16_"))))
def test_oracle_matches_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    orc, w = _oracle_for(str(g["voice"]))
    from oracle.make_golden import weights_digest
    assert weights_digest(w) == str(g["weights_sha256"]), "voice weights differ from the ones the golden was minted on"
    dump = {}
    o = orc.infer(g["ids"], g["scales"], g["eps_dp"] if "eps_dp" in g else None,
                  g["eps_z"] if "eps_z" in g else None, dump=dump, sid=int(g["sid"]) if "sid" in g else None)
    assert np.array_equal(dump["w_ceil"].numpy().astype(np.int32), g["w_ceil"])   # durations: exact
    assert o.shape == g["audio"].shape
    assert np.abs(dump["z"].numpy() - g["z"]).max() <= 1e-3
    assert np.abs(o - g["audio"]).max() <= 1e-3       # north_star tolerance, fp32 waveform


def test_oracle_matches_survey_anchors():
    """Frame counts / waveform statistics of all 7 en-us fixtures on the real voice."""
    lines = real_fixture_lines()
    if lines is None:
        pytest.skip("reference fixtures not staged")
    orc, _ = _oracle_for("real:test_voice")
    anchors = json.load(open(os.path.join(GOLDEN, "real_enus_anchors.json")))
    for a in anchors[1:4]:
        dump = {}
        o = orc.infer(lines[a["index"]]["phoneme_ids"], (0.0, 1.0, 0.0), dump=dump)
        assert len(o) == a["frames"] * 256
        assert dump["w_ceil"].numpy().astype(int).tolist() == a["w_ceil"]
        assert abs(float(np.abs(o).max()) - a["max_abs"]) < 1e-3
        assert abs(float(np.abs(o).mean()) - a["mean_abs"]) < 1e-4
        assert np.abs(o[:3] - np.asarray(a["first3"], np.float32)).max() < 1e-3


@pytest.mark.needs_reference
@pytest.mark.parametrize("tag,n_ph,scales", [("real:test_voice", 0, (0.667, 1.0, 0.8)),
                                             ("synthetic:tiny:1234", 40, (0.667, 1.3, 0.8)),
                                             ("synthetic:tiny-high:1234", 12, (0.3, 0.8, 1.0)),
                                             ("synthetic:tiny-ms:1234", 25, (0.667, 1.0, 0.8))])
def test_oracle_matches_reference_source(tag, n_ph, scales):
    from oracle import ref_bridge
    if not ref_bridge.available():
        pytest.skip("/root/reference not present (GPU box)")
    orc, w = _oracle_for(tag)
    net = ref_bridge.build_reference_model(orc.s, w)
    if tag.startswith("real:"):
        ids = real_fixture_lines()[3]["phoneme_ids"]
    else:
        ids = voicegen.benchmark_ids(n_ph, seed=17)
    rng = np.random.default_rng(11)
    eps_dp = rng.standard_normal((2, len(ids))).astype(np.float32)
    eps_z = rng.standard_normal((orc.s.inter, 6 * len(ids))).astype(np.float32)
    sid = 2 if orc.s.n_speakers > 1 else None             # multi-speaker: emb_g / dp.cond / WN cond_layer / dec.cond
    dump = {}
    o = orc.infer(ids, scales, eps_dp, eps_z, dump=dump, sid=sid)
    r = ref_bridge.reference_infer(net, ids, scales, eps_dp, eps_z, sid=sid)
    assert np.array_equal(dump["w_ceil"].numpy(), r["w_ceil"])
    assert np.abs(dump["z_p"].numpy() - r["z_p"]).max() < 1e-4
    assert np.abs(dump["z"].numpy() - r["z"]).max() < 1e-3
    assert np.abs(o - r["o"]).max() < 1e-3


def test_oracle_edge_cases():
    """Shortest legal input (BOS, PAD, EOS), single id, and a duration override."""
    orc, _ = _oracle_for("synthetic:tiny:1234")
    o = orc.infer([1, 0, 2], (0.667, 1.0, 0.8))
    assert len(o) % 256 == 0 and len(o) >= 256 and np.isfinite(o).all()
    o1 = orc.infer([5], (0.0, 1.0, 0.0))
    assert len(o1) >= 256
    o2 = orc.infer([1, 0, 7, 0, 2], (0.0, 1.0, 0.0), w_ceil_override=[1, 2, 3, 0, 1])
    assert len(o2) == 7 * 256
    # all-zero durations: clamp_min(sum, 1) gives exactly one frame (models.py:704)
    o3 = orc.infer([1, 0, 2], (0.0, 1.0, 0.0), w_ceil_override=[0, 0, 0])
    assert len(o3) == 256


def test_streaming_chunks_match_golden():
    """Chunk plan + oracle decoder against the fixture minted with the reference's own chunk loop."""
    from piper_b200 import streaming
    g = np.load(os.path.join(GOLDEN, "stream_tiny.npz"))
    orc, _ = _oracle_for(str(g["voice"]))
    z_p = orc.encode(g["ids"], g["scales"], g["eps_dp"], g["eps_z"])
    assert np.abs(z_p - g["z_p"]).max() <= 1e-4

    class _V:
        hop = orc.s.hop
    st = streaming.SpeechStreamer(_V(), int(g["chunk_size"]), int(g["chunk_padding"]), reference_quirks=True)
    pieces = list(st.chunk(z_p, decode=orc.decode))
    assert [len(p) for p in pieces] == g["piece_lens"].tolist()
    assert np.abs(np.concatenate(pieces) - g["audio"]).max() <= 1e-3
    # without the reference's stale right-trim the last chunk keeps its tail: total = frames * hop
    full = list(streaming.SpeechStreamer(_V(), 45, 10, reference_quirks=False).chunk(z_p, decode=orc.decode))
    assert sum(len(p) for p in full) == z_p.shape[1] * orc.s.hop


def test_chunk_plan_properties():
    from piper_b200.streaming import plan_chunks
    assert plan_chunks(65) == [(0, 65, 0, 0)]                       # too short to stream
    for n in (66, 90, 135, 136, 161, 500):
        plan = plan_chunks(n, 45, 10, reference_quirks=False)
        kept = sum((hi - lo) - tl - tr for lo, hi, tl, tr in plan)
        assert kept == n and plan[0][0] == 0 and plan[-1][1] == n
        assert all(hi - lo <= 45 + 20 for lo, hi, _, _ in plan)


@pytest.mark.needs_reference
def test_streaming_matches_reference_chunk_loop():
    from oracle import ref_bridge
    from piper_b200 import streaming
    if not ref_bridge.available():
        pytest.skip("/root/reference not present (GPU box)")
    orc, w = _oracle_for("synthetic:tiny:1234")
    net = ref_bridge.build_reference_model(orc.s, w)
    ids = voicegen.benchmark_ids(70, seed=5)
    z_p = orc.encode(ids, (0.667, 1.0, 0.8))
    ref = ref_bridge.reference_stream_chunks(net, z_p, z_p.shape[1], 45, 10)

    class _V:
        hop = orc.s.hop
    mine = list(streaming.SpeechStreamer(_V(), 45, 10, True).chunk(z_p, decode=orc.decode))
    assert [len(a) for a in mine] == [len(a) for a in ref]
    assert max(np.abs(a - b).max() for a, b in zip(mine, ref)) <= 1e-4


# ---- int16 peak normalisation (piper.cpp:411-431 / python_run/piper/util.py:5-12): the oracle pinned three ways

def _int16_golden():
    g = np.load(os.path.join(GOLDEN, "int16_cpp.npz"))
    return {k[:-3]: (g[k], g[k[:-3] + ".cpp"], g[k[:-3] + ".py"]) for k in g.files if k.endswith(".in")}


def test_int16_oracle_matches_reference_minted_golden():
    """tests/golden/int16_cpp.npz holds the outputs of the reference's own loops (compiled from piper.cpp:411-431)."""
    from oracle import int16_oracle
    from piper_b200 import host
    cases = _int16_golden()
    assert len(cases) >= 10
    for name, (a, cpp, py) in cases.items():
        assert np.array_equal(int16_oracle.float_to_int16_cpp(a), cpp), name
        assert np.array_equal(int16_oracle.float_to_int16_python(a), py), name
        assert np.array_equal(host.audio_float_to_int16(a), cpp), name            # the shim's Python mirror
        assert np.array_equal(host.audio_float_to_int16(a, "python"), py), name


def test_int16_oracle_matches_compiled_reference_loops():
    """oracle/_ref/libint16_ref.so = the reference's lines themselves (oracle/build_ref.py); travels to the GPU box."""
    from oracle import int16_oracle
    if int16_oracle.ref_lib() is None:
        pytest.skip("oracle/_ref/libint16_ref.so not built (python -m oracle.build_ref needs /root/reference)")
    rng = np.random.default_rng(11)
    for n, amp in ((1, 1.0), (7, 0.004), (1000, 0.5), (65536, 1.0), (4099, 3.0)):
        a = (amp * np.tanh(rng.standard_normal(n))).astype(np.float32) if amp <= 1 else (amp * rng.standard_normal(n)).astype(np.float32)
        assert np.array_equal(int16_oracle.float_to_int16_cpp(a), int16_oracle.float_to_int16_ref(a)), (n, amp)
    for name, (a, cpp, _) in _int16_golden().items():
        assert np.array_equal(int16_oracle.float_to_int16_ref(a), cpp), name


@pytest.mark.needs_reference
def test_int16_oracle_matches_reference_python_util():
    import importlib.util
    util_py = "/root/reference/src/python_run/piper/util.py"
    if not os.path.exists(util_py):
        pytest.skip("/root/reference not present (GPU box)")
    from oracle import int16_oracle
    spec = importlib.util.spec_from_file_location("ref_piper_util", util_py)
    util = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(util)
    rng = np.random.default_rng(12)
    for n, amp in ((3, 1.0), (500, 0.003), (30000, 0.7)):
        a = (amp * np.tanh(rng.standard_normal(n))).astype(np.float32)
        assert np.array_equal(int16_oracle.float_to_int16_python(a), util.audio_float_to_int16(a))
