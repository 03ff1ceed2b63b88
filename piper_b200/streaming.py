"""Chunked streaming synthesis: the reference's own notion of "streaming"
(/root/reference/src/python/piper_train/infer_onnx_streaming.py:32-39,76-124) over the engine's encode / decode split
(`pb200_encode` = VitsEncoder, `pb200_decode` = VitsDecoder of export_onnx_streaming.py:19-69).

The encoder runs once per utterance; flow + generator run on chunks of `chunk_size` frames with `chunk_padding` frames
of context on each side, whose audio is trimmed (`padding * hop` samples).  The halo is smaller than the decoder's true
receptive field, so chunked audio approximates full-utterance audio in the reference too: parity is
against the chunked reference, not against `synthesize`.

`reference_quirks=True` reproduces the reference loop exactly, including its stale `wav_end_pad`: the last chunk has no
right context, yet it is trimmed by the previous iteration's right-pad length (infer_onnx_streaming.py:88-108).  Pass
False to keep those samples.
"""
from __future__ import annotations

import math
from typing import Callable, Iterator, List, Optional, Tuple

import numpy as np


def plan_chunks(n_frames: int, chunk_size: int = 45, chunk_padding: int = 10, reference_quirks: bool = True
                ) -> List[Tuple[int, int, int, int]]:
    """-> [(first frame incl. left halo, one-past-last frame incl. right halo, samples-to-trim-left in frames,
    samples-to-trim-right in frames)] in playback order."""
    if n_frames <= chunk_size + 2 * chunk_padding:
        return [(0, n_frames, 0, 0)]
    bounds = [i * chunk_size for i in range(0, math.ceil(n_frames / chunk_size))] + [n_frames]
    chunks = [(bounds[i], bounds[i + 1]) for i in range(len(bounds) - 1)]
    plan = []
    stale_end = 0
    for idx, (a, b) in enumerate(chunks):
        left = right = 0
        lo, hi = a, b
        if idx > 0:
            pa, pb = chunks[idx - 1]
            left = min(chunk_padding, pb - pa)
            lo = a - left
        if idx + 1 < len(chunks):
            na, nb = chunks[idx + 1]
            right = min(chunk_padding, nb - na)
            hi = b + right
            stale_end = right
        trim_right = right if idx + 1 < len(chunks) else (stale_end if reference_quirks else 0)
        plan.append((lo, hi, left, trim_right))
    return plan


class SpeechStreamer:
    def __init__(self, voice, chunk_size: int = 45, chunk_padding: int = 10, reference_quirks: bool = True):
        self.voice = voice
        self.chunk_size = chunk_size
        self.chunk_padding = chunk_padding
        self.reference_quirks = reference_quirks

    def chunk(self, z_p: np.ndarray, decode: Optional[Callable[[np.ndarray], np.ndarray]] = None) -> Iterator[np.ndarray]:
        """z_p [inter][frames] -> fp32 audio pieces in playback order."""
        decode = decode or (lambda z: self.voice.decode(z)[0])
        hop = self.voice.hop
        for lo, hi, tl, tr in plan_chunks(z_p.shape[1], self.chunk_size, self.chunk_padding, self.reference_quirks):
            audio = decode(np.ascontiguousarray(z_p[:, lo:hi]))
            end = len(audio) - tr * hop
            yield audio[tl * hop:end]

    def stream(self, ids, scales=(0.667, 1.0, 0.8), seed: int = 0, eps_dp=None, eps_z=None) -> Iterator[np.ndarray]:
        z_p = self.voice.encode(ids, scales, eps_dp=eps_dp, eps_z=eps_z, seed=seed)
        for piece in self.chunk(z_p):
            if len(piece):
                yield piece

    def stream_int16_bytes(self, ids, scales=(0.667, 1.0, 0.8), seed: int = 0) -> Iterator[bytes]:
        """What the reference's SpeechStreamer.stream yields: per-chunk peak-normalised int16 PCM."""
        from .host import audio_float_to_int16
        for piece in self.stream(ids, scales, seed):
            yield audio_float_to_int16(piece, variant="python").tobytes()
