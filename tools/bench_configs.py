"""Side measurements for the other BASELINE.json configs (not the driver's bench line): one JSON line each.
  config 2: medium, batch = 1 latency / RTF                (also in bench.py's `batch1`)
  config 4: high architecture, 8 utterances per GPU (the per-GPU share of batch 64 over 8 GPUs)
  config 5: generator only, z ~ N(0,1) [B,192,256], B in {1, 8, 32, 128}, medium and high decoders
usage: python tools/bench_configs.py > gpurun_out/configs.jsonl
"""
import json, os, statistics, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from piper_b200 import engine, voicegen

SCALES = (0.667, 1.0, 0.8)
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
# generator-only algorithmic bytes for 256 frames: activations + weights once per launch
GEN_BYTES = {"medium": (191.82e6, 6.65e6), "high": (1037.50e6, 57.31e6)}
GEN_FLOP = {"medium": 11.614e9, "high": 157.416e9}

def med(f, n=7):
    xs = []
    for _ in range(n):
        xs.append(f())
    return statistics.median(xs)

for arch in ("medium", "high"):
    v = engine.Voice(voicegen.cached_voice(arch), 0)
    ids1 = [voicegen.benchmark_ids(128, seed=1234)]
    for _ in range(3):
        v.synthesize_batch(ids1, SCALES, seed=1, copy=False)
    def one():
        t = time.perf_counter(); flat, counts, _ = v.synthesize_batch(ids1, SCALES, seed=1, copy=False); one.n = int(counts.sum()); return time.perf_counter() - t
    lat = med(one)
    print(json.dumps({"config": f"{arch} batch=1, 128 phonemes", "latency_ms": lat * 1e3, "samples": one.n, "rtf": lat / (one.n / 22050.0),
                      "samples_per_s": one.n / lat}), flush=True)
    if arch == "high":
        ids8 = [voicegen.benchmark_ids(128, seed=1234 + b) for b in range(8)]
        v.stage(ids8, SCALES, seed=4242)
        for _ in range(3):
            v.run_staged()
        tot_ms, n = 0.0, 0
        for _ in range(5):
            s, ms = v.run_staged(); tot_ms += ms; n += s
        print(json.dumps({"config": "high architecture, 8 x 128-phoneme utterances on one GPU (per-GPU share of config 4)",
                          "samples_per_s": n / (tot_ms * 1e-3), "ms_per_step": tot_ms / 5,
                          "stage_ms": dict(zip(["enc", "dp", "sync", "flow", "gen"], v.stage_times()))}), flush=True)
    for B in (1, 8, 32, 128):
        z = np.random.default_rng(1236).standard_normal((B, 192, 256)).astype(np.float32)
        for _ in range(2):
            v.vocode(z)
        ms = med(lambda: (v.vocode(z), v.stage_times()[4])[1], 5)
        act, w = GEN_BYTES[arch]
        gbs = (act * B + w) / (ms * 1e-3) / 1e9
        print(json.dumps({"config": f"generator only ({arch}), z[{B},192,256] -> [{B},65536]", "device_ms": ms,
                          "samples_per_s": B * 65536 / (ms * 1e-3), "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / PEAK,
                          "tflops_algorithmic": GEN_FLOP[arch] * B / (ms * 1e-3) / 1e12}), flush=True)
    v.close()
