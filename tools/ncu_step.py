"""Tiny driver for ncu captures: N staged steps of the bench workload (no CPU baseline, no extras).
usage: python tools/ncu_step.py [steps] [batch] [arch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from piper_b200 import engine, voicegen
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 32
arch = sys.argv[3] if len(sys.argv) > 3 else "medium"
v = engine.Voice(voicegen.cached_voice(arch), 0)
ids = [voicegen.benchmark_ids(128, seed=1234 + b) for b in range(batch)]
v.stage(ids, (0.667, 1.0, 0.8), seed=4242)
for s in range(steps):
    n, ms = v.run_staged()
    print(f"step {s}: {n} samples, {ms:.3f} ms, launches so far {engine.launch_count()}", flush=True)
