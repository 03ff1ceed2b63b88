"""Build oracle/_ref/ from the reference sources WHERE THEY LIE (no reference source is copied into the repository).

    python -m oracle.build_ref          # needs /root/reference; outputs only into oracle/_ref/ (git-ignored)

* libint16_ref.so — the reference's own int16 peak-normalisation loops, /root/reference/src/cpp/piper.cpp:411-431.
  piper.cpp as a whole needs onnxruntime, spdlog and piper-phonemize (absent), so the recipe cuts exactly those
  dependency-free lines out of the file at build time (located by the comments that bracket them, and checked against
  the expected line numbers) into oracle/_ref/int16_body.inc and compiles them inside oracle/int16_harness.cpp with g++.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SRC = "/root/reference/src/cpp/piper.cpp"
BEGIN, END = "// Get max audio value for scaling", "// Clean up"


def build_int16() -> str | None:
    if not os.path.exists(SRC):
        return None
    lines = open(SRC).read().splitlines()
    b = next(i for i, l in enumerate(lines) if l.strip() == BEGIN)
    e = next(i for i, l in enumerate(lines) if i > b and l.strip() == END)
    assert (b + 1, e) == (410, 432), f"piper.cpp moved: block found at lines {b + 1}-{e}"
    os.makedirs(OUT, exist_ok=True)
    inc = os.path.join(OUT, "int16_body.inc")
    with open(inc, "w") as f:
        f.write("\n".join(lines[b:e]) + "\n")
    lib = os.path.join(OUT, "libint16_ref.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", HERE, os.path.join(HERE, "int16_harness.cpp"),
                    "-o", lib], check=True)
    return lib


def main() -> int:
    lib = build_int16()
    print("built" if lib else "skipped (no /root/reference)", lib or "")
    return 0


if __name__ == "__main__":
    sys.exit(main())
