"""CPU fp32 restatement of piper's VITS inference graph (phoneme ids -> waveform).

TEST INFRASTRUCTURE (see oracle/__init__.py): the checker the CUDA engine is compared
against, and the "port" CPU baseline of bench.py.  Batch = 1 only — that is the only
mode any reference caller uses (`/root/reference/src/cpp/piper.cpp:352`,
`/root/reference/src/python_run/piper/voice.py:158`) and parity is defined per
utterance.  With B = 1 every sequence mask of the reference is all-ones, so masks do
not appear below.

Restated from (all under /root/reference/src/python/piper_train/vits/):
  SynthesizerTrn.infer          models.py:681-722
  TextEncoder.forward           models.py:198-209    Encoder       attentions.py:60-74
  MultiHeadAttention.attention  attentions.py:225-272 (+ rel-pos helpers :274-348)
  FFN.forward                   attentions.py:386-407  LayerNorm   modules.py:23-26
  StochasticDurationPredictor   models.py:63-70,108-117
  DDSConv / ConvFlow            modules.py:117-129 / 496-527
  rational-quadratic spline     transforms.py:50-98,101-191 (inverse branch)
  Flip / ElementwiseAffine      modules.py:384-409
  generate_path (as a gather)   commons.py:116-129
  ResidualCouplingLayer / WN    modules.py:447-466 / 184-209
  Generator / ResBlock1/2       models.py:348-368 ; modules.py:301-314,355-364

Noise is an explicit input (the graph's two RandomNormalLike nodes are unseeded):
`eps_dp [2,T]` for the duration flows and `eps_z [inter, >=T']` for the prior sample.

Parity status: pinned against the reference's own PyTorch source run in the build
container (tests/test_oracle_vs_reference.py, tests/golden/*.npz minted by
oracle/make_golden.py).  The reference holds no numeric golden vectors of its own
(its only test asserts WAV size >= 10 kB, src/cpp/test.cpp:52-55).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from .voice_loader import VoiceSpec, ConvAttr


def _t(w: Dict[str, np.ndarray], name: str) -> torch.Tensor:
    v = w[name]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))


class Oracle:
    def __init__(self, spec: VoiceSpec, weights: Dict[str, np.ndarray], attrs: Dict[str, ConvAttr]):
        self.s = spec
        self.w = {k: _t(weights, k) for k in weights}
        self.a = attrs

    # ------------------------------------------------------------------ helpers
    def conv(self, x, name, dilation=1, pad=0, groups=1):
        """x [C,T] -> [Co,T']  (Conv1d with weight `name.weight`, optional bias)."""
        b = self.w.get(name + ".bias")
        return F.conv1d(x[None], self.w[name + ".weight"], b, dilation=dilation, padding=pad,
                        groups=groups)[0]

    def layer_norm(self, x, prefix):
        g, b = self.w[prefix + ".gamma"], self.w[prefix + ".beta"]
        return F.layer_norm(x.t(), (x.shape[0],), g, b, 1e-5).t()

    # ------------------------------------------------------------- text encoder
    def attention(self, x, l):
        s = self.s
        p = f"enc_p.encoder.attn_layers.{l}"
        H, T = x.shape
        nh, dk, win = s.n_heads, s.hidden // s.n_heads, s.window
        q = self.conv(x, p + ".conv_q").view(nh, dk, T).transpose(1, 2)   # [h,T,dk]
        k = self.conv(x, p + ".conv_k").view(nh, dk, T).transpose(1, 2)
        v = self.conv(x, p + ".conv_v").view(nh, dk, T).transpose(1, 2)
        qs = q / math.sqrt(dk)
        scores = qs @ k.transpose(1, 2)                                   # [h,T,T]
        ek, ev = self.w[p + ".emb_rel_k"][0], self.w[p + ".emb_rel_v"][0]  # [2w+1,dk] (shared by heads)
        rel = qs @ ek.t()                                                 # [h,T,2w+1]
        i = torch.arange(T)
        for r in range(-win, win + 1):
            src = i[(i + r >= 0) & (i + r < T)]
            scores[:, src, src + r] += rel[:, src, r + win]
        pattn = torch.softmax(scores, dim=-1)
        out = pattn @ v                                                   # [h,T,dk]
        for r in range(-win, win + 1):
            src = i[(i + r >= 0) & (i + r < T)]
            out[:, src] += pattn[:, src, src + r][..., None] * ev[r + win]
        out = out.transpose(1, 2).reshape(H, T)
        return self.conv(out, p + ".conv_o")

    def text_encoder(self, ids: torch.Tensor):
        s = self.s
        x = (self.w["enc_p.emb.weight"][ids] * math.sqrt(s.hidden)).t().contiguous()   # [H,T]
        kp = s.ffn_kernel
        for l in range(s.n_layers):
            y = self.attention(x, l)
            x = self.layer_norm(x + y, f"enc_p.encoder.norm_layers_1.{l}")
            f = f"enc_p.encoder.ffn_layers.{l}"
            h = F.pad(x, ((kp - 1) // 2, kp // 2))
            h = torch.relu(self.conv(h, f + ".conv_1"))
            h = F.pad(h, ((kp - 1) // 2, kp // 2))
            y = self.conv(h, f + ".conv_2")
            x = self.layer_norm(x + y, f"enc_p.encoder.norm_layers_2.{l}")
        stats = self.conv(x, "enc_p.proj")
        return x, stats[: s.inter], stats[s.inter:]

    # ------------------------------------------------------ duration predictor
    def dds_conv(self, x, prefix, g=None):
        if g is not None:
            x = x + g
        for i in range(self.s.dds_layers):
            name = f"{prefix}.convs_sep.{i}"
            k = self.w[name + ".weight"].shape[2]
            d = k ** i
            y = self.conv(x, name, dilation=d, pad=(k * d - d) // 2, groups=x.shape[0])
            y = F.gelu(self.layer_norm(y, f"{prefix}.norms_1.{i}"))
            y = self.conv(y, f"{prefix}.convs_1x1.{i}")
            y = F.gelu(self.layer_norm(y, f"{prefix}.norms_2.{i}"))
            x = x + y
        return x

    def rqs_inverse(self, x, uw, uh, ud, bound=5.0):
        """x [T]; uw,uh [T,nb]; ud [T,nb-1].  Branch-free form of transforms.py:50-191."""
        nb = uw.shape[-1]
        const = float(np.log(np.exp(1 - 1e-3) - 1))
        ud = F.pad(ud, (1, 1), value=const)
        widths = 1e-3 + (1 - 1e-3 * nb) * torch.softmax(uw, -1)
        cw = F.pad(torch.cumsum(widths, -1), (1, 0)) * (2 * bound) - bound
        cw[..., 0], cw[..., -1] = -bound, bound
        widths = cw[..., 1:] - cw[..., :-1]
        deriv = 1e-3 + F.softplus(ud)
        heights = 1e-3 + (1 - 1e-3 * nb) * torch.softmax(uh, -1)
        ch = F.pad(torch.cumsum(heights, -1), (1, 0)) * (2 * bound) - bound
        ch[..., 0], ch[..., -1] = -bound, bound
        heights = ch[..., 1:] - ch[..., :-1]
        loc = ch.clone()
        loc[..., -1] += 1e-6
        inside = (x >= -bound) & (x <= bound)
        xc = torch.where(inside, x, torch.zeros_like(x))
        idx = ((xc[..., None] >= loc).sum(-1) - 1).clamp(0, nb - 1)[..., None]
        g = lambda t: t.gather(-1, idx)[..., 0]
        in_cw, in_w, in_ch, in_h = g(cw), g(widths), g(ch), g(heights)
        in_delta = g(heights / widths)
        d0, d1 = g(deriv), g(deriv[..., 1:])
        t = (xc - in_ch) * (d0 + d1 - 2 * in_delta)
        a = t + in_h * (in_delta - d0)
        b = in_h * d0 - t
        c = -in_delta * (xc - in_ch)
        disc = b.pow(2) - 4 * a * c
        root = (2 * c) / (-b - torch.sqrt(disc))
        out = root * in_w + in_cw
        return torch.where(inside, out, x)

    def duration_logw(self, x, eps_dp, noise_w, g=None):
        s = self.s
        c = self.conv(x, "dp.pre")
        if g is not None:                                   # models.py:66-68
            c = c + self.conv(g, "dp.cond")
        c = self.dds_conv(c, "dp.convs")
        c = self.conv(c, "dp.proj")
        z = eps_dp * noise_w                                             # [2,T]
        nb = s.spline_bins
        for f in s.dp_flows:
            z = z.flip(0)
            x0, x1 = z[0:1], z[1]
            h = self.conv(x0, f"dp.flows.{f}.pre")
            h = self.dds_conv(h, f"dp.flows.{f}.convs", g=c)
            h = self.conv(h, f"dp.flows.{f}.proj").t()                   # [T, 3nb-1]
            uw = h[:, :nb] / math.sqrt(s.hidden)
            uh = h[:, nb:2 * nb] / math.sqrt(s.hidden)
            ud = h[:, 2 * nb:]
            x1 = self.rqs_inverse(x1, uw, uh, ud)
            z = torch.stack([x0[0], x1])
        z = z.flip(0)
        z = (z - self.w["dp.flows.0.m"]) * torch.exp(-self.w["dp.flows.0.logs"])
        return z[0]

    # ------------------------------------------------------------------- flow
    def wn(self, h, prefix, g=None):
        s = self.s
        H = h.shape[0]
        out = torch.zeros_like(h)
        k = s.wn_kernel
        gc = self.conv(g, f"{prefix}.cond_layer") if g is not None else None      # modules.py:188-197
        for i in range(s.wn_layers):
            d = s.wn_dilation_rate ** i
            a = self.conv(h, f"{prefix}.in_layers.{i}", dilation=d, pad=(k * d - d) // 2)
            if gc is not None:
                a = a + gc[i * 2 * H:(i + 1) * 2 * H]
            acts = torch.tanh(a[:H]) * torch.sigmoid(a[H:])
            rs = self.conv(acts, f"{prefix}.res_skip_layers.{i}")
            if i < s.wn_layers - 1:
                h = h + rs[:H]
                out = out + rs[H:]
            else:
                out = out + rs
        return out

    def flow_reverse(self, z, g=None):
        half = self.s.inter // 2
        for f in self.s.flow_layers:
            z = z.flip(0)
            x0, x1 = z[:half], z[half:]
            h = self.conv(x0, f"flow.flows.{f}.pre")
            h = self.wn(h, f"flow.flows.{f}.enc", g)
            m = self.conv(h, f"flow.flows.{f}.post")
            z = torch.cat([x0, x1 - m], 0)
        return z

    # -------------------------------------------------------------- generator
    def generator(self, z, dump: Optional[dict] = None, g=None):
        s = self.s
        x = self.conv(z, "dec.conv_pre", pad=3)
        if g is not None:                                   # models.py:350-351
            x = x + self.conv(g, "dec.cond")
        nk = len(s.rb_kernels)
        for i, (u, k, p) in enumerate(zip(s.up_rates, s.up_kernels, s.up_pads)):
            x = F.leaky_relu(x, 0.1)
            x = F.conv_transpose1d(x[None], self.w[f"dec.ups.{i}.weight"], self.w.get(f"dec.ups.{i}.bias"),
                                   stride=u, padding=p)[0]
            if dump is not None:
                dump[f"up{i}"] = x
            xs = None
            for j in range(nk):
                rb = f"dec.resblocks.{i * nk + j}"
                kk = s.rb_kernels[j]
                y = x
                for c, d in enumerate(s.rb_dilations[j]):
                    if s.resblock == 1:
                        t = self.conv(F.leaky_relu(y, 0.1), f"{rb}.convs1.{c}", dilation=d, pad=d * (kk - 1) // 2)
                        t = self.conv(F.leaky_relu(t, 0.1), f"{rb}.convs2.{c}", pad=(kk - 1) // 2)
                    else:
                        t = self.conv(F.leaky_relu(y, 0.1), f"{rb}.convs.{c}", dilation=d, pad=d * (kk - 1) // 2)
                    y = t + y
                xs = y if xs is None else xs + y
            x = xs / nk
            if dump is not None:
                dump[f"stage{i}"] = x
        x = F.leaky_relu(x)                         # default slope 0.01 (models.py:364)
        x = self.conv(x, "dec.conv_post", pad=3)
        return torch.tanh(x)[0]

    # ------------------------------------------------- streaming split (export_onnx_streaming.py:19-69)
    @torch.no_grad()
    def encode(self, ids, scales, eps_dp=None, eps_z=None):
        """VitsEncoder.forward: everything before the flow -> z_p [inter, T'] (np.ndarray)."""
        dump = {}
        self.infer(ids, scales, eps_dp, eps_z, dump=dump, stop_before_flow=True)
        return dump["z_p"].numpy()

    @torch.no_grad()
    def decode(self, z_p):
        """VitsDecoder.forward: flow reverse + generator on z_p [inter, frames] -> waveform."""
        z = self.flow_reverse(torch.as_tensor(np.asarray(z_p), dtype=torch.float32))
        return self.generator(z).numpy()

    # ------------------------------------------------------------------ infer
    @torch.no_grad()
    def infer(self, ids, scales, eps_dp=None, eps_z=None, w_ceil_override=None, dump: Optional[dict] = None,
              stop_before_flow: bool = False, sid: Optional[int] = None):
        """ids int64 [T]; scales = (noise_scale, length_scale, noise_w).
        eps_dp [2,T] / eps_z [inter, >=T'] default to zeros (deterministic graph).
        Returns fp32 waveform [T' * hop] (np.ndarray)."""
        s = self.s
        ids = torch.as_tensor(np.asarray(ids), dtype=torch.long)
        T = ids.numel()
        noise_scale, length_scale, noise_w = (float(v) for v in scales)
        x, m_p, logs_p = self.text_encoder(ids)
        g = None
        if s.n_speakers > 1:                                # models.py:692-696
            g = self.w["emb_g.weight"][int(sid or 0)][:, None]
        if eps_dp is None:
            eps_dp = torch.zeros(2, T)
        eps_dp = torch.as_tensor(np.asarray(eps_dp), dtype=torch.float32)
        logw = self.duration_logw(x, eps_dp, noise_w, g)
        w = torch.exp(logw) * length_scale
        w_ceil = torch.ceil(w)
        if w_ceil_override is not None:
            w_ceil = torch.as_tensor(np.asarray(w_ceil_override), dtype=torch.float32)
        total = int(w_ceil.sum().item())
        y_len = max(total, 1)
        cum = torch.cumsum(w_ceil, 0)
        j = torch.arange(y_len, dtype=torch.float32)
        idx = torch.searchsorted(cum, j, right=True).clamp(max=T - 1)
        valid = (j < float(total)).float()
        m_e = m_p[:, idx] * valid
        logs_e = logs_p[:, idx] * valid
        if eps_z is None:
            eps_z = torch.zeros(s.inter, y_len)
        eps_z = torch.as_tensor(np.asarray(eps_z), dtype=torch.float32)[:, :y_len]
        z_p = m_e + eps_z * torch.exp(logs_e) * noise_scale
        if stop_before_flow:
            dump.update(z_p=z_p, w_ceil=w_ceil)
            return None
        z = self.flow_reverse(z_p, g)
        o = self.generator(z, dump, g)
        if dump is not None:
            dump.update(x=x, m_p=m_p, logs_p=logs_p, logw=logw, w_ceil=w_ceil, z_p=z_p, z=z, o=o)
        return o.numpy()
