"""ORACLE (test infrastructure, never shipped): minimal ONNX protobuf reader.

No `onnx` / `onnxruntime` package exists in this image, so the oracle reads the
reference's model files (`src/real_time_vad/models/silero_vad*.onnx`) with a
hand-rolled protobuf wire decoder.  Only the message fields that the two Silero
graphs use are decoded (field numbers follow the public onnx.proto3 schema).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

# ---------------------------------------------------------------- wire level


def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not (b & 0x80):
            return result, pos
        shift += 7


def _signed64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _fields(buf: bytes):
    """Yield (field_number, wire_type, value) for one message body."""
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield fno, wt, v


def _packed_varints(v, wt) -> List[int]:
    if wt == 0:
        return [_signed64(v)]
    out = []
    pos = 0
    while pos < len(v):
        x, pos = _varint(v, pos)
        out.append(_signed64(x))
    return out


# ---------------------------------------------------------------- messages

_DTYPES = {1: np.float32, 2: np.uint8, 3: np.int8, 5: np.int16, 6: np.int32,
           7: np.int64, 9: np.bool_, 10: np.float16, 11: np.float64}


def parse_tensor(buf: bytes) -> Tuple[str, np.ndarray]:
    dims: List[int] = []
    dtype = 1
    name = ""
    raw = None
    f32: List[float] = []
    i32: List[int] = []
    i64: List[int] = []
    f64: List[float] = []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            dims += _packed_varints(v, wt)
        elif fno == 2:
            dtype = v
        elif fno == 4:
            if wt == 2:
                f32 += list(struct.unpack(f"<{len(v) // 4}f", v))
            else:
                f32.append(struct.unpack("<f", v)[0])
        elif fno == 5:
            i32 += _packed_varints(v, wt)
        elif fno == 7:
            i64 += _packed_varints(v, wt)
        elif fno == 8:
            name = v.decode()
        elif fno == 9:
            raw = bytes(v)
        elif fno == 10:
            if wt == 2:
                f64 += list(struct.unpack(f"<{len(v) // 8}d", v))
            else:
                f64.append(struct.unpack("<d", v)[0])
    np_dt = _DTYPES[dtype]
    if raw is not None:
        arr = np.frombuffer(raw, dtype=np_dt).copy()
    elif f32:
        arr = np.asarray(f32, dtype=np_dt)
    elif i64:
        arr = np.asarray(i64, dtype=np_dt)
    elif i32:
        arr = np.asarray(i32, dtype=np_dt)
    elif f64:
        arr = np.asarray(f64, dtype=np_dt)
    else:
        arr = np.zeros(0, dtype=np_dt)
    arr = arr.reshape(dims) if dims else (arr.reshape(()) if arr.size == 1 else arr)
    return name, arr


@dataclass
class Node:
    op: str
    name: str
    inputs: List[str]
    outputs: List[str]
    attrs: Dict[str, Any] = field(default_factory=dict)


@dataclass
class Graph:
    name: str
    nodes: List[Node]
    initializers: Dict[str, np.ndarray]
    inputs: List[str]
    outputs: List[str]


def _parse_attr(buf: bytes) -> Tuple[str, Any]:
    name = ""
    atype = 0
    f = i = s = t = g = None
    floats: List[float] = []
    ints: List[int] = []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            name = v.decode()
        elif fno == 2:
            f = struct.unpack("<f", v)[0]
        elif fno == 3:
            i = _signed64(v)
        elif fno == 4:
            s = bytes(v)
        elif fno == 5:
            t = parse_tensor(v)[1]
        elif fno == 6:
            g = parse_graph(v)
        elif fno == 7:
            if wt == 2:
                floats += list(struct.unpack(f"<{len(v) // 4}f", v))
            else:
                floats.append(struct.unpack("<f", v)[0])
        elif fno == 8:
            ints += _packed_varints(v, wt)
        elif fno == 20:
            atype = v
    # AttributeProto.AttributeType: FLOAT=1 INT=2 STRING=3 TENSOR=4 GRAPH=5 FLOATS=6 INTS=7
    if atype == 1:
        return name, f
    if atype == 2:
        return name, i
    if atype == 3:
        return name, s
    if atype == 4:
        return name, t
    if atype == 5:
        return name, g
    if atype == 6:
        return name, floats
    if atype == 7:
        return name, ints
    for cand in (g, t, s, i, f):
        if cand is not None:
            return name, cand
    return name, ints or floats


def _parse_node(buf: bytes) -> Node:
    ins: List[str] = []
    outs: List[str] = []
    name = op = ""
    attrs: Dict[str, Any] = {}
    for fno, wt, v in _fields(buf):
        if fno == 1:
            ins.append(v.decode())
        elif fno == 2:
            outs.append(v.decode())
        elif fno == 3:
            name = v.decode()
        elif fno == 4:
            op = v.decode()
        elif fno == 5:
            k, val = _parse_attr(v)
            attrs[k] = val
    return Node(op, name, ins, outs, attrs)


def _value_info_name(buf: bytes) -> str:
    for fno, wt, v in _fields(buf):
        if fno == 1:
            return v.decode()
    return ""


def parse_graph(buf: bytes) -> Graph:
    nodes: List[Node] = []
    inits: Dict[str, np.ndarray] = {}
    ins: List[str] = []
    outs: List[str] = []
    name = ""
    for fno, wt, v in _fields(buf):
        if fno == 1:
            nodes.append(_parse_node(v))
        elif fno == 2:
            name = v.decode()
        elif fno == 5:
            k, arr = parse_tensor(v)
            inits[k] = arr
        elif fno == 11:
            ins.append(_value_info_name(v))
        elif fno == 12:
            outs.append(_value_info_name(v))
    return Graph(name, nodes, inits, ins, outs)


@dataclass
class Model:
    graph: Graph
    producer: str
    opset: int


def load_model(path: str) -> Model:
    with open(path, "rb") as fh:
        buf = fh.read()
    graph: Optional[Graph] = None
    producer = ""
    opset = 0
    for fno, wt, v in _fields(buf):
        if fno == 7:
            graph = parse_graph(v)
        elif fno == 2:
            producer = v.decode()
        elif fno == 8:  # OperatorSetIdProto {domain=1, version=2}
            dom, ver = "", 0
            for f2, w2, v2 in _fields(v):
                if f2 == 1:
                    dom = v2.decode()
                elif f2 == 2:
                    ver = v2
            if dom in ("", "ai.onnx"):
                opset = ver
    if graph is None:
        raise ValueError(f"no graph in {path}")
    return Model(graph, producer, opset)
