"""Mint tests/golden/int16_cpp.npz from the reference's own loops (oracle/_ref/libint16_ref.so = piper.cpp:411-431
compiled by oracle/build_ref.py) and from the reference's Python twin (python_run/piper/util.py:5-12, imported
read-only).  Build container only:  python -m oracle.build_ref && python -m oracle.make_golden_int16
"""
from __future__ import annotations

import importlib.util
import os

import numpy as np

from oracle import int16_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UTIL = "/root/reference/src/python_run/piper/util.py"


def cases():
    rng = np.random.default_rng(20260922)
    c = {
        "zeros": np.zeros(64, np.float32),
        "below_floor": np.full(96, 0.001, np.float32) * rng.choice([-1, 1], 96).astype(np.float32),   # peak < 0.01 -> floor
        "at_floor": np.array([0.01, -0.01, 0.005, -0.0099999], np.float32),
        "single": np.array([0.37], np.float32),
        "neg_peak": np.concatenate([[-0.9], rng.uniform(-0.5, 0.5, 255)]).astype(np.float32),       # peak is negative
        "speech_like": (0.3 * rng.standard_normal(8192)).astype(np.float32).clip(-0.98, 0.98),
        "tanh_out": np.tanh(rng.standard_normal(4096) * 2).astype(np.float32),                       # saturating waveform
        "tiny_denormal": (rng.standard_normal(128) * 1e-39).astype(np.float32),
        "ramp": np.linspace(-1, 1, 4097, dtype=np.float32),
    }
    # products that land within an ulp of an integer / of the clamp edges
    k = np.arange(-32767, 32768, 257, dtype=np.float32)
    c["near_integers"] = np.concatenate([k / np.float32(32767.0), np.nextafter(k / np.float32(32767.0), np.float32(2)),
                                         np.nextafter(k / np.float32(32767.0), np.float32(-2))]).astype(np.float32)
    return c


def main():
    spec = importlib.util.spec_from_file_location("ref_piper_util", UTIL)
    util = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(util)
    out = {}
    for name, a in cases().items():
        out[f"{name}.in"] = a
        out[f"{name}.cpp"] = int16_oracle.float_to_int16_ref(a)
        out[f"{name}.py"] = util.audio_float_to_int16(a)
        print(f"{name}: n={a.size} cpp[:4]={out[f'{name}.cpp'][:4]} differs_from_python={int((out[f'{name}.cpp'] != out[f'{name}.py']).sum())}")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "int16_cpp.npz"), **out)


if __name__ == "__main__":
    main()
