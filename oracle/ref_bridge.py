"""Run the reference's OWN PyTorch model source (read-only import from /root/reference).

TEST INFRASTRUCTURE, build container only: `/root/reference` does not exist on the GPU
box, so this module is used (a) by tests marked `needs_reference` to validate the
restatement in oracle/vits_oracle.py, and (b) by oracle/make_golden.py to mint the
fixtures committed under tests/golden/.  Recipe follows SURVEY.md Appendix C.
"""
from __future__ import annotations

import os
import sys
import types
import warnings
from typing import Dict

import numpy as np
import torch

REF_PY = "/root/reference/src/python"


def available() -> bool:
    return os.path.isdir(os.path.join(REF_PY, "piper_train", "vits"))


def _import_models():
    if REF_PY not in sys.path:
        sys.path.insert(0, REF_PY)
    stub = "piper_train.vits.monotonic_align"
    if stub not in sys.modules:          # training-only Cython module, never used by infer
        m = types.ModuleType(stub)
        m.maximum_path = None
        sys.modules[stub] = m
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from piper_train.vits import models  # type: ignore
    return models


def build_reference_model(spec, weights: Dict[str, np.ndarray]):
    """Instantiate SynthesizerTrn (models.py:520-615) and fill it with `weights`."""
    models = _import_models()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = models.SynthesizerTrn(
            n_vocab=spec.n_vocab, spec_channels=513, segment_size=32,
            inter_channels=spec.inter, hidden_channels=spec.hidden, filter_channels=spec.filter,
            n_heads=spec.n_heads, n_layers=spec.n_layers, kernel_size=spec.ffn_kernel, p_dropout=0.1,
            resblock=str(spec.resblock), resblock_kernel_sizes=tuple(spec.rb_kernels),
            resblock_dilation_sizes=tuple(tuple(d) for d in spec.rb_dilations),
            upsample_rates=tuple(spec.up_rates), upsample_initial_channel=spec.up_initial,
            upsample_kernel_sizes=tuple(spec.up_kernels), n_speakers=getattr(spec, "n_speakers", 1),
            gin_channels=getattr(spec, "gin", 0), use_sdp=True)
        net.eval()
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            net.dec.remove_weight_norm()
        for f in net.flow.flows:
            if hasattr(f, "enc"):
                f.enc.remove_weight_norm()
    sd = net.state_dict()
    missing = [k for k in sd if k not in weights
               and not k.startswith(("enc_q.", "dp.post_", "dp.flows.1."))]
    unexpected = [k for k in weights if k not in sd]
    if missing or unexpected:
        raise RuntimeError(f"state-dict mismatch: missing={missing[:5]} unexpected={unexpected[:5]}")
    with torch.no_grad():
        for k, v in weights.items():
            sd[k].copy_(torch.from_numpy(np.ascontiguousarray(v)).reshape(sd[k].shape))
    return net


@torch.no_grad()
def reference_infer(net, ids, scales, eps_dp=None, eps_z=None, sid=None):
    """Call the reference's `SynthesizerTrn.infer` (models.py:681-722) with injected noise.

    torch.randn / torch.randn_like are patched for the duration of the call so that the
    two in-graph noise draws (models.py:111 and :718) return the supplied tensors.
    Returns dict(o, w_ceil, z, z_p, y_len)."""
    ids_t = torch.as_tensor(np.asarray(ids), dtype=torch.long)[None]
    lens = torch.tensor([ids_t.shape[1]], dtype=torch.long)
    T = ids_t.shape[1]
    real_randn, real_randn_like = torch.randn, torch.randn_like

    def fake_randn(*size, **kw):
        shape = tuple(size[0]) if len(size) == 1 and not isinstance(size[0], int) else tuple(size)
        if shape == (1, 2, T):
            if eps_dp is None:
                return torch.zeros(shape)
            return torch.as_tensor(np.asarray(eps_dp), dtype=torch.float32).reshape(shape)
        return real_randn(*size, **kw)

    def fake_randn_like(t, **kw):
        if eps_z is None:
            return torch.zeros_like(t)
        e = torch.as_tensor(np.asarray(eps_z), dtype=torch.float32)
        return e[:, : t.shape[2]].reshape(1, t.shape[1], t.shape[2]).clone()

    torch.randn, torch.randn_like = fake_randn, fake_randn_like
    try:
        o, attn, y_mask, (z, z_p, m_p, logs_p) = net.infer(
            ids_t, lens, sid=None if sid is None else torch.tensor([int(sid)]), noise_scale=float(scales[0]),
            length_scale=float(scales[1]), noise_scale_w=float(scales[2]))
    finally:
        torch.randn, torch.randn_like = real_randn, real_randn_like
    w_ceil = attn[0, 0].sum(0)
    return dict(o=o[0, 0].numpy(), w_ceil=w_ceil.numpy(), z=z[0].numpy(), z_p=z_p[0].numpy(),
                y_len=int(y_mask.sum().item()))


def reference_stream_chunks(net, z_p, y_len, chunk_size=45, chunk_padding=10):
    """Run the reference's OWN chunking loop (`SpeechStreamer.chunk`,
    /root/reference/src/python/piper_train/infer_onnx_streaming.py:76-108) with its decoder call replaced by the
    reference PyTorch decoder (flow reverse + generator = VitsDecoder, export_onnx_streaming.py:61-69).
    onnxruntime is absent, so the module is imported with a stub in its place; only `chunk` is exercised."""
    import importlib
    if REF_PY not in sys.path:
        sys.path.insert(0, REF_PY)
    for name in ("onnxruntime",):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    _import_models()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mod = importlib.import_module("piper_train.infer_onnx_streaming")
    streamer = mod.SpeechStreamer.__new__(mod.SpeechStreamer)
    streamer.chunk_size, streamer.chunk_padding, streamer.sample_rate = chunk_size, chunk_padding, 22050

    def decoder_infer(z, y_mask, g=None):
        with torch.no_grad():
            zt, mt = torch.from_numpy(np.ascontiguousarray(z)), torch.from_numpy(np.ascontiguousarray(y_mask))
            zz = net.flow(zt, mt, g=None, reverse=True)
            return net.dec(zz * mt).squeeze().numpy()

    streamer.decoder_infer = decoder_infer
    z = np.asarray(z_p, np.float32)[None]
    y_mask = np.ones((1, 1, z.shape[2]), np.float32)
    out = streamer.chunk([z, y_mask])
    if isinstance(out, np.ndarray):          # "too short to stream": chunk() returns the audio itself
        return [out]
    return [np.asarray(a) for a in out]
