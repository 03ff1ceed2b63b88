"""Host-side logic either side of the hot path, mirroring the reference's own host code.

* `phonemes_to_ids`  — id layout the C++ host feeds `synthesize` with: BOS, PAD, (phoneme ids, PAD)*, EOS.
  The implementation lives in piper-phonemize (external, fetched unpinned as master.zip by
  /root/reference/CMakeLists.txt:63-71; call site src/cpp/piper.cpp:555; id constants
  src/cpp/piper.hpp:44-47).  Restated here from its published behaviour and pinned on the reference's
  own pre-phonemized fixtures etc/test_sentences/test_en-us.jsonl (7/7 lines reproduce).
  The Python runtime's variant (src/python_run/piper/voice.py:72-87) omits the PAD after BOS; it is
  available as `layout="python"`.
* `audio_float_to_int16` — peak normalisation of src/cpp/piper.cpp:411-431 (truncating cast), with the
  numpy twin of src/python_run/piper/util.py:5-12 selectable.
* `wav_header` — the 44-byte header of src/cpp/wavfile.hpp:6-38.
* `VoiceConfig` — the fields of `<voice>.onnx.json` the C++ host parses (src/cpp/piper.cpp:47-214).
"""
from __future__ import annotations

import json
import struct
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

PAD, BOS, EOS = "_", "^", "$"
ID_PAD, ID_BOS, ID_EOS = 0, 1, 2
MAX_WAV_VALUE = 32767.0


def phonemes_to_ids(phonemes: Iterable[str], id_map: Dict[str, Sequence[int]], layout: str = "cpp",
                    missing: Optional[Dict[str, int]] = None) -> List[int]:
    pad = list(id_map.get(PAD, [ID_PAD]))
    ids: List[int] = list(id_map.get(BOS, [ID_BOS]))
    if layout == "cpp":
        ids.extend(pad)
    elif layout != "python":
        raise ValueError(f"unknown layout {layout!r}")
    for p in phonemes:
        if p not in id_map:
            if missing is not None:
                missing[p] = missing.get(p, 0) + 1
            continue
        ids.extend(id_map[p])
        ids.extend(pad)
    ids.extend(id_map.get(EOS, [ID_EOS]))
    return ids


def audio_float_to_int16(audio: np.ndarray, variant: str = "cpp") -> np.ndarray:
    audio = np.asarray(audio, np.float32)
    peak = np.float32(max(0.01, float(np.max(np.abs(audio))) if audio.size else 0.0))
    scale = np.float32(MAX_WAV_VALUE) / np.float32(max(np.float32(0.01), peak))
    if variant == "cpp":      # clamp to [int16 min, int16 max], static_cast truncates toward zero
        v = np.clip(audio * scale, np.float32(-32768.0), np.float32(32767.0))
        return np.trunc(v).astype(np.int16)
    if variant == "python":   # util.py clips symmetrically to +-32767
        v = np.clip(audio * (MAX_WAV_VALUE / max(0.01, float(np.max(np.abs(audio))))), -MAX_WAV_VALUE, MAX_WAV_VALUE)
        return v.astype("int16")
    raise ValueError(variant)


def wav_header(sample_rate: int, sample_width: int, channels: int, num_samples: int) -> bytes:
    data_size = num_samples * sample_width * channels
    return struct.pack("<4sI4s4sIHHIIHH4sI", b"RIFF", data_size + 44 - 8, b"WAVE", b"fmt ", 16, 1, channels,
                       sample_rate, sample_rate * sample_width * channels, sample_width * channels, 16, b"data",
                       data_size)


def write_wav(path: str, audio_i16: np.ndarray, sample_rate: int) -> None:
    a = np.ascontiguousarray(audio_i16, dtype="<i2")
    with open(path, "wb") as f:
        f.write(wav_header(sample_rate, 2, 1, a.size))
        f.write(a.tobytes())


@dataclass
class VoiceConfig:
    sample_rate: int = 22050
    noise_scale: float = 0.667
    length_scale: float = 1.0
    noise_w: float = 0.8
    num_speakers: int = 1
    espeak_voice: str = "en-us"
    phoneme_type: str = "espeak"
    phoneme_id_map: Dict[str, List[int]] = field(default_factory=dict)
    speaker_id_map: Dict[str, int] = field(default_factory=dict)
    phoneme_silence: Dict[str, float] = field(default_factory=dict)

    @staticmethod
    def load(path: str) -> "VoiceConfig":
        with open(path, "r", encoding="utf-8") as f:
            root = json.load(f)
        c = VoiceConfig()
        c.sample_rate = int(root.get("audio", {}).get("sample_rate", c.sample_rate))
        inf = root.get("inference", {})
        c.noise_scale = float(inf.get("noise_scale", c.noise_scale))
        c.length_scale = float(inf.get("length_scale", c.length_scale))
        c.noise_w = float(inf.get("noise_w", c.noise_w))
        c.phoneme_silence = {k: float(v) for k, v in inf.get("phoneme_silence", {}).items()}
        c.num_speakers = int(root.get("num_speakers", 1))
        c.espeak_voice = root.get("espeak", {}).get("voice", c.espeak_voice)
        c.phoneme_type = root.get("phoneme_type", c.phoneme_type)
        c.phoneme_id_map = {k: list(v) for k, v in root.get("phoneme_id_map", {}).items()}
        c.speaker_id_map = dict(root.get("speaker_id_map", {}))
        return c

    @property
    def scales(self) -> Tuple[float, float, float]:
        return (self.noise_scale, self.length_scale, self.noise_w)


def shard_utterances(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Length-balanced assignment of utterance indices to ranks: longest first onto the
    currently lightest rank.  Deterministic; every rank computes the same plan, no collective needed."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    loads = [0] * world_size
    plan: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        plan[r].append(i)
        loads[r] += int(lengths[i])
    for p in plan:
        p.sort()
    return plan
