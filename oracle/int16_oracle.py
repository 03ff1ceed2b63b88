"""CPU oracle of piper's float -> int16 peak normalisation — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Two reference statements exist and they differ on purpose:

* C++ host, /root/reference/src/cpp/piper.cpp:411-431 — `maxAudioValue` starts at 0.01f and is raised by
  `abs(audio[i])` in a float loop; `audioScale = MAX_WAV_VALUE / max(0.01f, maxAudioValue)` in float (MAX_WAV_VALUE =
  32767.0f, piper.cpp:29); each sample is `static_cast<int16_t>(clamp(audio[i] * audioScale, -32768.f, 32767.f))`, i.e. a
  float product, clamped, truncated toward zero.  This is what `pb200_synthesize_int16` must reproduce bit for bit.
* Python runtime, /root/reference/src/python_run/piper/util.py:5-12 — scale computed in float64 from a Python float
  `max(0.01, np.max(np.abs(audio)))`, the float32 array multiplied by that Python scalar (numpy keeps float32),
  clipped to +-32767 and cast with `astype("int16")` (truncation).

Pinned by tests/test_oracle.py against (a) util.py imported read-only from /root/reference, (b) oracle/_ref/libint16_ref.so
— the lines piper.cpp:411-431 themselves, extracted at build time by oracle/build_ref.py and compiled by g++ — and (c)
tests/golden/int16_cpp.npz minted from (b) by oracle/make_golden_int16.py.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

MAX_WAV_VALUE = np.float32(32767.0)          # piper.cpp:29
REF_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libint16_ref.so")


def float_to_int16_cpp(audio: np.ndarray) -> np.ndarray:
    """piper.cpp:411-431 in float32 arithmetic."""
    a = np.ascontiguousarray(audio, np.float32).reshape(-1)
    peak = np.float32(0.01)
    if a.size:
        peak = np.maximum(peak, np.abs(a).max()).astype(np.float32)     # running `if (v > max) max = v` == max
    scale = np.float32(MAX_WAV_VALUE / np.maximum(np.float32(0.01), peak))
    v = a * scale                                                      # float32 product, as `audio[i] * audioScale`
    v = np.clip(v, np.float32(-32768.0), np.float32(32767.0))
    return np.trunc(v).astype(np.int16)                                # static_cast truncates toward zero


def float_to_int16_python(audio: np.ndarray) -> np.ndarray:
    """python_run/piper/util.py:5-12."""
    a = np.asarray(audio)
    norm = a * (32767.0 / max(0.01, float(np.max(np.abs(a)))))
    return np.clip(norm, -32767.0, 32767.0).astype("int16")


def ref_lib():
    """The compiled extract of piper.cpp:411-431 (oracle/build_ref.py), or None when it has not been built."""
    if not os.path.exists(REF_LIB):
        return None
    lib = ctypes.CDLL(REF_LIB)
    lib.ref_float_to_int16.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.c_int64, ctypes.POINTER(ctypes.c_int16)]
    lib.ref_float_to_int16.restype = None
    return lib


def float_to_int16_ref(audio: np.ndarray) -> np.ndarray:
    """Run the reference's own loop (compiled from its source)."""
    lib = ref_lib()
    if lib is None:
        raise FileNotFoundError(REF_LIB)
    a = np.ascontiguousarray(audio, np.float32).reshape(-1)
    out = np.empty(a.size, np.int16)
    lib.ref_float_to_int16(a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), a.size,
                           out.ctypes.data_as(ctypes.POINTER(ctypes.c_int16)))
    return out
