"""Minimal ONNX protobuf wire reader/writer (no `onnx` / protobuf dependency).

Only what a piper voice file needs: the graph's initializers (TensorProto) and the
Conv / ConvTranspose nodes with their attributes, which is all the engine's loader
(`csrc/onnx_reader.cc`, the C++ twin of this file) consumes.  The wire layout is the
public protobuf encoding of onnx.proto (ModelProto.graph = 7, GraphProto.node = 1,
GraphProto.initializer = 5, TensorProto.{dims=1,data_type=2,name=8,raw_data=9}, ...).

The writer emits files with the same container layout as the reference exporter
(`/root/reference/src/python/piper_train/export_onnx.py:88-101`): initializers with
`raw_data`, Conv nodes referencing weight / bias initializers by name.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np

# ONNX TensorProto.DataType
DT_FLOAT, DT_INT64 = 1, 7
_NP = {DT_FLOAT: np.float32, DT_INT64: np.int64}


def _varint(buf: memoryview, pos: int) -> Tuple[int, int]:
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _fields(buf: memoryview) -> Iterator[Tuple[int, int, object]]:
    """Yield (field_number, wire_type, value) for one message body."""
    pos, end = 0, len(buf)
    while pos < end:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
            yield fno, wt, v
        elif wt == 1:
            yield fno, wt, bytes(buf[pos:pos + 8])
            pos += 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            if pos + n > end:
                raise ValueError("truncated length-delimited protobuf field")
            yield fno, wt, buf[pos:pos + n]
            pos += n
        elif wt == 5:
            yield fno, wt, bytes(buf[pos:pos + 4])
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")


def _signed(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


@dataclass
class Node:
    op_type: str = ""
    name: str = ""
    inputs: List[str] = field(default_factory=list)
    outputs: List[str] = field(default_factory=list)
    ints: Dict[str, List[int]] = field(default_factory=dict)    # ints / single int attrs
    floats: Dict[str, List[float]] = field(default_factory=dict)


@dataclass
class Model:
    producer: str = ""
    ir_version: int = 0
    opset: int = 0
    initializers: Dict[str, np.ndarray] = field(default_factory=dict)
    init_order: List[str] = field(default_factory=list)
    nodes: List[Node] = field(default_factory=list)
    inputs: List[str] = field(default_factory=list)
    outputs: List[str] = field(default_factory=list)


def _parse_tensor(buf: memoryview) -> Tuple[str, np.ndarray]:
    dims: List[int] = []
    dtype = DT_FLOAT
    name = ""
    raw: Optional[memoryview] = None
    f32: List[float] = []
    i64: List[int] = []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            if wt == 0:
                dims.append(_signed(v))
            else:  # packed
                p = 0
                while p < len(v):
                    d, p = _varint(v, p)
                    dims.append(_signed(d))
        elif fno == 2:
            dtype = v
        elif fno == 8:
            name = bytes(v).decode()
        elif fno == 9:
            raw = v
        elif fno == 4:  # float_data
            if wt == 2:
                f32.extend(np.frombuffer(bytes(v), dtype="<f4").tolist())
            else:
                f32.append(struct.unpack("<f", v)[0])
        elif fno == 7:  # int64_data
            if wt == 2:
                p = 0
                while p < len(v):
                    d, p = _varint(v, p)
                    i64.append(_signed(d))
            else:
                i64.append(_signed(v))
    if dtype not in _NP:
        return name, np.zeros(0, np.float32)
    if raw is not None:
        arr = np.frombuffer(bytes(raw), dtype=np.dtype(_NP[dtype]).newbyteorder("<"))
    elif dtype == DT_FLOAT:
        arr = np.asarray(f32, np.float32)
    else:
        arr = np.asarray(i64, np.int64)
    return name, arr.reshape(dims).astype(_NP[dtype], copy=True)


def _parse_attr(buf: memoryview, node: Node) -> None:
    name = ""
    ints: List[int] = []
    floats: List[float] = []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            name = bytes(v).decode()
        elif fno == 3:
            ints.append(_signed(v))
        elif fno == 8:
            if wt == 2:
                p = 0
                while p < len(v):
                    d, p = _varint(v, p)
                    ints.append(_signed(d))
            else:
                ints.append(_signed(v))
        elif fno == 2:
            floats.append(struct.unpack("<f", v)[0])
        elif fno == 7:
            if wt == 2:
                floats.extend(np.frombuffer(bytes(v), dtype="<f4").tolist())
            else:
                floats.append(struct.unpack("<f", v)[0])
    if ints:
        node.ints[name] = ints
    if floats:
        node.floats[name] = floats


def _parse_node(buf: memoryview) -> Node:
    n = Node()
    for fno, wt, v in _fields(buf):
        if fno == 1:
            n.inputs.append(bytes(v).decode())
        elif fno == 2:
            n.outputs.append(bytes(v).decode())
        elif fno == 3:
            n.name = bytes(v).decode()
        elif fno == 4:
            n.op_type = bytes(v).decode()
        elif fno == 5:
            _parse_attr(v, n)
    return n


def _value_info_name(buf: memoryview) -> str:
    for fno, wt, v in _fields(buf):
        if fno == 1:
            return bytes(v).decode()
    return ""


def load(path: str) -> Model:
    with open(path, "rb") as f:
        data = memoryview(f.read())
    m = Model()
    for fno, wt, v in _fields(data):
        if fno == 1:
            m.ir_version = v
        elif fno == 2:
            m.producer = bytes(v).decode()
        elif fno == 8:
            for f2, _, v2 in _fields(v):
                if f2 == 2:
                    m.opset = v2
        elif fno == 7:
            for f2, w2, v2 in _fields(v):
                if f2 == 1:
                    m.nodes.append(_parse_node(v2))
                elif f2 == 5:
                    name, arr = _parse_tensor(v2)
                    m.initializers[name] = arr
                    m.init_order.append(name)
                elif f2 == 11:
                    m.inputs.append(_value_info_name(v2))
                elif f2 == 12:
                    m.outputs.append(_value_info_name(v2))
    return m


# --------------------------------------------------------------------------- writer

def _enc_varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _key(fno: int, wt: int) -> bytes:
    return _enc_varint((fno << 3) | wt)


def _ld(fno: int, payload: bytes) -> bytes:
    return _key(fno, 2) + _enc_varint(len(payload)) + payload


def _vi(fno: int, v: int) -> bytes:
    return _key(fno, 0) + _enc_varint(v)


def _enc_tensor(name: str, arr: np.ndarray) -> bytes:
    if arr.dtype == np.float32:
        dt = DT_FLOAT
    elif arr.dtype == np.int64:
        dt = DT_INT64
    else:
        raise TypeError(arr.dtype)
    out = bytearray()
    for d in arr.shape:
        out += _vi(1, int(d))
    out += _vi(2, dt)
    out += _ld(8, name.encode())
    out += _ld(9, np.ascontiguousarray(arr).astype(arr.dtype.newbyteorder("<")).tobytes())
    return bytes(out)


def _enc_attr_ints(name: str, vals: List[int]) -> bytes:
    out = bytearray(_ld(1, name.encode()))
    for v in vals:
        out += _vi(8, int(v))
    out += _vi(20, 7)  # AttributeType.INTS
    return bytes(out)


def _enc_attr_int(name: str, val: int) -> bytes:
    return _ld(1, name.encode()) + _vi(3, int(val)) + _vi(20, 2)


def _enc_node(n: Node) -> bytes:
    out = bytearray()
    for s in n.inputs:
        out += _ld(1, s.encode())
    for s in n.outputs:
        out += _ld(2, s.encode())
    out += _ld(3, n.name.encode())
    out += _ld(4, n.op_type.encode())
    for k, v in n.ints.items():
        if k == "group":
            out += _ld(5, _enc_attr_int(k, v[0]))
        else:
            out += _ld(5, _enc_attr_ints(k, v))
    return bytes(out)


def save(path: str, m: Model) -> None:
    g = bytearray()
    for n in m.nodes:
        g += _ld(1, _enc_node(n))
    g += _ld(2, b"torch-jit-export")
    for name in m.init_order:
        g += _ld(5, _enc_tensor(name, m.initializers[name]))
    for name in m.inputs:
        g += _ld(11, _ld(1, name.encode()))
    for name in m.outputs:
        g += _ld(12, _ld(1, name.encode()))
    out = bytearray()
    out += _vi(1, m.ir_version or 8)
    out += _ld(2, (m.producer or "piper_b200.voicegen").encode())
    out += _ld(7, bytes(g))
    out += _ld(8, _ld(1, b"") + _vi(2, m.opset or 15))
    with open(path, "wb") as f:
        f.write(bytes(out))
