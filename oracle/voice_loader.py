"""Oracle-side voice loader: `.onnx` initializers -> canonical weight dict + VoiceSpec.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The product loader is the C++ twin in
`piper_b200/csrc/voice.cc`; both restate the same naming rules, which come from how
`/root/reference/src/python/piper_train/export_onnx.py:51-101` serialises
`SynthesizerTrn` (`/root/reference/src/python/piper_train/vits/models.py:520-615`):

* the embedding table is the initializer called `sid` (4 input_names, sid=None),
* weight-normed flow convs are constant-folded into anonymous `onnx::Conv_*`
  initializers -> recovered through the Conv node's *bias* input name,
* `dp.flows.0.logs` is absent; `exp(-logs)` survives as the [2,1] Mul operand that
  follows `Sub(., dp.flows.0.m)`.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

from piper_b200 import onnx_wire


@dataclass
class ConvAttr:
    kernel: int
    dilation: int = 1
    stride: int = 1
    pad: int = 0
    groups: int = 1


@dataclass
class VoiceSpec:
    n_vocab: int = 0
    hidden: int = 0          # H (text encoder / SDP / flow-WN hidden)
    inter: int = 0           # inter_channels (flow / generator input)
    filter: int = 0          # FFN filter channels
    n_heads: int = 0
    n_layers: int = 0
    window: int = 0
    ffn_kernel: int = 0
    dds_layers: int = 3
    dp_flows: List[int] = field(default_factory=list)     # ConvFlow indices in execution order (e.g. 7,5,3)
    spline_bins: int = 10
    flow_layers: List[int] = field(default_factory=list)  # coupling indices in execution order (6,4,2,0)
    wn_layers: int = 0
    wn_kernel: int = 0
    wn_dilation_rate: int = 1
    resblock: int = 2         # 1 or 2
    up_rates: List[int] = field(default_factory=list)
    up_kernels: List[int] = field(default_factory=list)
    up_pads: List[int] = field(default_factory=list)
    up_initial: int = 0
    rb_kernels: List[int] = field(default_factory=list)
    rb_dilations: List[List[int]] = field(default_factory=list)
    hop: int = 0
    n_speakers: int = 1
    gin: int = 0             # speaker-embedding width (0 = single speaker)


def canonicalize(model: onnx_wire.Model) -> Tuple[Dict[str, np.ndarray], Dict[str, ConvAttr]]:
    """Return (weights by reference state-dict name, conv attrs by weight name)."""
    init = model.initializers
    w: Dict[str, np.ndarray] = {}
    attrs: Dict[str, ConvAttr] = {}
    anonymous = set()
    for n in model.nodes:
        if n.op_type not in ("Conv", "ConvTranspose") or len(n.inputs) < 2:
            continue
        wname = n.inputs[1]
        if wname not in init:
            continue
        canon = wname
        if wname.startswith("onnx::") and len(n.inputs) >= 3 and n.inputs[2].endswith(".bias"):
            canon = n.inputs[2][: -len(".bias")] + ".weight"
            anonymous.add(wname)
        arr = init[wname]
        k = int(n.ints.get("kernel_shape", [arr.shape[-1]])[0])
        a = ConvAttr(kernel=k,
                     dilation=int(n.ints.get("dilations", [1])[0]),
                     stride=int(n.ints.get("strides", [1])[0]),
                     pad=int(n.ints.get("pads", [0, 0])[0]),
                     groups=int(n.ints.get("group", [1])[0]))
        attrs[canon] = a
        w[canon] = arr
    for name, arr in init.items():
        if name.startswith("onnx::") or name in anonymous:
            continue
        if arr.dtype != np.float32:
            continue
        w.setdefault(name, arr)
    if "enc_p.emb.weight" not in w:
        if "sid" in init and init["sid"].ndim == 2:
            w["enc_p.emb.weight"] = init["sid"]
            w.pop("sid", None)
        else:
            raise ValueError("embedding table not found (neither enc_p.emb.weight nor 2-D `sid`)")
    if "dp.flows.0.logs" not in w:
        # (z - m) * exp(-logs): Sub(z, dp.flows.0.m) -> Mul(., onnx::Mul_NNNN [2,1])
        producers = {}
        for n in model.nodes:
            if n.op_type == "Sub" and "dp.flows.0.m" in n.inputs:
                producers[n.outputs[0]] = n
        found = None
        for n in model.nodes:
            if n.op_type == "Mul" and any(i in producers for i in n.inputs):
                for i in n.inputs:
                    if i in init and init[i].shape == (2, 1):
                        found = init[i]
        if found is None:
            raise ValueError("dp.flows.0 exp(-logs) constant not found")
        w["dp.flows.0.logs"] = (-np.log(found)).astype(np.float32)
    return w, attrs


def _indices(w: Dict[str, np.ndarray], pattern: str) -> List[int]:
    rx = re.compile(pattern)
    return sorted({int(m.group(1)) for k in w for m in [rx.match(k)] if m})


def infer_spec(w: Dict[str, np.ndarray], attrs: Dict[str, ConvAttr]) -> VoiceSpec:
    s = VoiceSpec()
    emb = w["enc_p.emb.weight"]
    s.n_vocab, s.hidden = int(emb.shape[0]), int(emb.shape[1])
    relk = w["enc_p.encoder.attn_layers.0.emb_rel_k"]
    dk = int(relk.shape[2])
    s.n_heads = s.hidden // dk
    s.window = (int(relk.shape[1]) - 1) // 2
    s.n_layers = len(_indices(w, r"enc_p\.encoder\.attn_layers\.(\d+)\.conv_q\.weight"))
    c1 = w["enc_p.encoder.ffn_layers.0.conv_1.weight"]
    s.filter, s.ffn_kernel = int(c1.shape[0]), int(c1.shape[2])
    s.inter = int(w["enc_p.proj.weight"].shape[0]) // 2
    s.dds_layers = len(_indices(w, r"dp\.convs\.convs_sep\.(\d+)\.weight"))
    cf = _indices(w, r"dp\.flows\.(\d+)\.pre\.weight")
    if len(cf) > 1 and min(cf) == 1:      # self.flows[1] is dropped by the reverse pass (models.py:110) even if a file keeps it
        cf = [i for i in cf if i != 1]
    s.dp_flows = sorted(cf, reverse=True)
    s.spline_bins = (int(w[f"dp.flows.{cf[0]}.proj.weight"].shape[0]) + 1) // 3
    fl = _indices(w, r"flow\.flows\.(\d+)\.pre\.weight")
    s.flow_layers = sorted(fl, reverse=True)
    s.wn_layers = len(_indices(w, rf"flow\.flows\.{fl[0]}\.enc\.in_layers\.(\d+)\.weight"))
    k0 = f"flow.flows.{fl[0]}.enc.in_layers.0.weight"
    s.wn_kernel = int(w[k0].shape[2])
    if s.wn_layers > 1:
        s.wn_dilation_rate = attrs[f"flow.flows.{fl[0]}.enc.in_layers.1.weight"].dilation
    ups = _indices(w, r"dec\.ups\.(\d+)\.weight")
    s.up_initial = int(w["dec.conv_pre.weight"].shape[0])
    for i in ups:
        a = attrs[f"dec.ups.{i}.weight"]
        s.up_rates.append(a.stride)
        s.up_kernels.append(a.kernel)
        s.up_pads.append(a.pad)
    s.hop = int(np.prod(s.up_rates))
    if "emb_g.weight" in w:
        s.n_speakers, s.gin = int(w["emb_g.weight"].shape[0]), int(w["emb_g.weight"].shape[1])
    s.resblock = 1 if any(k.startswith("dec.resblocks.0.convs1.") for k in w) else 2
    n_rb = len(_indices(w, r"dec\.resblocks\.(\d+)\."))
    per_stage = n_rb // len(ups)
    for j in range(per_stage):
        if s.resblock == 1:
            names = [f"dec.resblocks.{j}.convs1.{i}.weight"
                     for i in _indices(w, rf"dec\.resblocks\.{j}\.convs1\.(\d+)\.weight")]
        else:
            names = [f"dec.resblocks.{j}.convs.{i}.weight"
                     for i in _indices(w, rf"dec\.resblocks\.{j}\.convs\.(\d+)\.weight")]
        s.rb_kernels.append(int(w[names[0]].shape[2]))
        s.rb_dilations.append([attrs[n].dilation for n in names])
    return s


def load_voice(path: str):
    """-> (VoiceSpec, {name: np.ndarray fp32}, {weight name: ConvAttr})"""
    model = onnx_wire.load(path)
    w, attrs = canonicalize(model)
    return infer_spec(w, attrs), w, attrs
