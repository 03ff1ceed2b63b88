"""Per-stage GPU-vs-oracle diagnostics (developer tool; run under gpurun).  Writes gpurun_out/diag.txt."""
import json, os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from piper_b200 import engine, voicegen
from oracle.voice_loader import load_voice
from oracle.vits_oracle import Oracle

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", "diag.txt"), "w")
def P(*a):
    s = " ".join(str(x) for x in a)
    print(s); out.write(s + "\n"); out.flush()

def diag(path, ids, scales, seed, tag):
    P(f"==== {tag}: {os.path.basename(path)} ids={len(ids)} scales={scales}")
    spec, w, attrs = load_voice(path)
    orc = Oracle(spec, w, attrs)
    rng = np.random.default_rng(seed)
    eps_dp = rng.standard_normal((2, len(ids))).astype(np.float32)
    eps_z = rng.standard_normal((spec.inter, 4 * len(ids) + 64)).astype(np.float32)
    d = {}
    t = time.time(); ref = orc.infer(ids, scales, eps_dp, eps_z, dump=d); tc = time.time() - t
    v = engine.Voice(path, 0)
    v.set_debug(True)
    try:
        audio, sec = v.synthesize(ids, scales, eps_dp, eps_z)
    except Exception as e:
        P("  synthesize FAILED:", e); traceback.print_exc(); return
    def cmp(name, got, exp):
        exp = np.asarray(exp, np.float32)
        if got.shape != exp.shape:
            P(f"  {name:8s} SHAPE gpu {got.shape} oracle {exp.shape}"); return
        e = np.abs(got - exp)
        P(f"  {name:8s} shape {str(got.shape):16s} max|err| {e.max():.3e}  rms(ref) {np.sqrt((exp**2).mean()):.3e}  argmax {np.unravel_index(e.argmax(), e.shape)}")
    cmp("x", v.tap("x"), d["x"].numpy())
    st = v.tap("stats")
    cmp("m_p", st[:spec.inter], d["m_p"].numpy()); cmp("logs_p", st[spec.inter:], d["logs_p"].numpy())
    cmp("logw", v.tap("logw")[0], d["logw"].numpy())
    cum = v.tap("cum")[0]
    wc = np.diff(np.concatenate([[0], cum]))
    P("  w_ceil equal:", np.array_equal(wc, d["w_ceil"].numpy()), "frames gpu", int(cum[-1]), "oracle", int(d["w_ceil"].sum()))
    for k in ["z_p", "z"] + [f"up{i}" for i in range(len(spec.up_rates))] + [f"stage{i}" for i in range(len(spec.up_rates))]:
        try:
            cmp(k, v.tap(k), d[k].numpy())
        except Exception as e:
            P(f"  {k}: {e}")
    if audio.shape == ref.shape:
        P(f"  AUDIO    n={len(audio)} max|err| {np.abs(audio-ref).max():.3e} rms(ref) {np.sqrt((ref**2).mean()):.3f}  gpu {sec*1e3:.2f} ms  cpu-oracle {tc*1e3:.0f} ms")
    else:
        P(f"  AUDIO SHAPE gpu {audio.shape} oracle {ref.shape}")
    v.set_debug(False)
    for _ in range(3):
        audio, sec = v.synthesize(ids, scales, eps_dp, eps_z)
    P(f"  warm synthesize: {sec*1e3:.3f} ms for {len(audio)} samples -> {len(audio)/sec/1e6:.2f} Msamples/s; stage ms {v.stage_times()}")
    v.close()

real = os.path.join(ROOT, "oracle", "_ref", "voice", "test_voice.onnx")
try:
    diag(voicegen.cached_voice("tiny"), voicegen.benchmark_ids(20), (0.667, 1.0, 0.8), 1, "tiny")
    diag(voicegen.cached_voice("tiny-high"), voicegen.benchmark_ids(20), (0.667, 1.0, 0.8), 2, "tiny-high")
    if os.path.exists(real):
        lines = [json.loads(l) for l in open(os.path.join(os.path.dirname(real), "test_en-us.jsonl"))]
        diag(real, lines[1]["phoneme_ids"], (0.0, 1.0, 0.0), 3, "real-voice det")
        diag(real, lines[2]["phoneme_ids"], (0.667, 1.0, 0.8), 4, "real-voice noise")
    diag(voicegen.cached_voice("medium"), voicegen.benchmark_ids(128), (0.667, 1.0, 0.8), 5, "medium")
    diag(voicegen.cached_voice("high"), voicegen.benchmark_ids(128), (0.667, 1.0, 0.8), 6, "high")
except Exception:
    P(traceback.format_exc())
P("launches:", engine.launch_count())
