"""Per-tap max-abs error of the engine against the oracle for one utterance / one batch (developer diagnostic).
usage: [env switches] python tools/tap_errors.py arch n_phonemes [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from piper_b200 import engine, voicegen
from oracle.voice_loader import load_voice
from oracle.vits_oracle import Oracle

arch, n_ph = sys.argv[1], int(sys.argv[2])
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 1
path = voicegen.cached_voice(arch) if arch != "real" else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "voice", "test_voice.onnx")
spec, w, attrs = load_voice(path)
orc = Oracle(spec, w, attrs)
ids = [voicegen.benchmark_ids(n_ph, seed=3 + b, n_vocab=min(256, spec.n_vocab)) for b in range(batch)]
rng = np.random.default_rng(77)
eps_dp = [rng.standard_normal((2, len(i))).astype(np.float32) for i in ids]
eps_z = rng.standard_normal((batch, spec.inter, 6 * len(ids[0]) + 16)).astype(np.float32)
v = engine.Voice(path, 0)
v.set_debug(True)
wavs, _ = v.synthesize_batch(ids, (0.667, 1.0, 0.8), eps_dp=eps_dp, eps_z=eps_z)
keys = ["x", "stats", "logw", "cum", "z_p", "z"] + [f"up{i}" for i in range(len(spec.up_rates))] + [f"stage{i}" for i in range(len(spec.up_rates))]
for b in sorted(set([0, batch - 1])):
    dump = {}
    ref = orc.infer(ids[b], (0.667, 1.0, 0.8), eps_dp[b], eps_z[b], dump=dump)
    out = []
    for k in keys:
        try:
            t = v.tap(k, b)
        except Exception as e:
            out.append(f"{k}=n/a"); continue
        if k == "stats":
            r = np.concatenate([dump["m_p"].numpy(), dump["logs_p"].numpy()])
        elif k == "cum":
            r = np.cumsum(dump["w_ceil"].numpy())[None].astype(np.float32)
        elif k == "logw":
            r = dump["logw"].numpy()[None] if dump["logw"].ndim == 1 else dump["logw"].numpy()
        elif k in dump:
            r = dump[k].numpy()
        else:
            out.append(f"{k}=noref"); continue
        if t.shape != r.shape:
            out.append(f"{k}=SHAPE{t.shape}vs{r.shape}")
        else:
            out.append(f"{k}={np.abs(t - r).max():.2e}")
    wave = f"wave={np.abs(wavs[b] - ref).max():.2e}" if wavs[b].shape == ref.shape else f"wave=SHAPE{wavs[b].shape}vs{ref.shape}"
    print(f"{arch}/{n_ph}/B{batch} item {b}: " + " ".join(out) + " " + wave, flush=True)
v.close()
