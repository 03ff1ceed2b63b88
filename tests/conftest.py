import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_ROOT = "/root/reference"
REAL_VOICE_CANDIDATES = [
    os.path.join(REF_ROOT, "etc", "test_voice.onnx"),
    os.path.join(ROOT, "oracle", "_ref", "voice", "test_voice.onnx"),
]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "needs_reference: imports the reference source from /root/reference")


def real_voice_path():
    for p in REAL_VOICE_CANDIDATES:
        if os.path.exists(p):
            return p
    return None


def real_fixture_lines():
    import json
    for p in (os.path.join(REF_ROOT, "etc", "test_sentences", "test_en-us.jsonl"),
              os.path.join(ROOT, "oracle", "_ref", "voice", "test_en-us.jsonl")):
        if os.path.exists(p):
            return [json.loads(l) for l in open(p)]
    return None


@pytest.fixture(scope="session")
def lib_built():
    """Build the C-ABI library once per session (nvcc cross-compiles without a GPU)."""
    from piper_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.load()
