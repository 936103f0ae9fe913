"""Weight extraction from the Silero `.onnx` files shipped in `real_time_vad/models/`.

The reference hands the model file to onnxruntime
(/root/reference/src/real_time_vad/core/silero_model.py:321-325); this engine only
needs the tensors, so it walks the protobuf wire format directly (no `onnx`
package exists in the target image) and collects every float tensor it finds --
graph initializers and `Constant` node values, at any subgraph depth -- keyed by
name.  `canonical_blob()` then lays the tensors of one model branch out in the
order `cvad_create()` expects (include/cutter_vad_b200.h).
"""
from __future__ import annotations

import struct
from pathlib import Path
from typing import Dict, Iterator, List, Tuple, Union

import numpy as np

# protobuf field numbers (onnx.proto3)
_MODEL_GRAPH = 7
_GRAPH_NODE, _GRAPH_INIT = 1, 5
_NODE_OUTPUT, _NODE_OPTYPE, _NODE_ATTR = 2, 4, 5
_ATTR_T, _ATTR_G, _ATTR_GRAPHS = 5, 6, 11
_T_DIMS, _T_DTYPE, _T_FLOATS, _T_NAME, _T_RAW = 1, 2, 4, 8, 9
_FLOAT = 1


def _read_varint(b: memoryview, i: int) -> Tuple[int, int]:
    out = shift = 0
    while True:
        byte = b[i]
        i += 1
        out |= (byte & 0x7F) << shift
        if byte < 0x80:
            return out, i
        shift += 7


def _walk(b: memoryview) -> Iterator[Tuple[int, int, Union[int, memoryview]]]:
    i, n = 0, len(b)
    while i < n:
        tag, i = _read_varint(b, i)
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            val, i = _read_varint(b, i)
        elif wire == 2:
            ln, i = _read_varint(b, i)
            val = b[i:i + ln]
            i += ln
        elif wire == 5:
            val = b[i:i + 4]
            i += 4
        elif wire == 1:
            val = b[i:i + 8]
            i += 8
        else:
            raise ValueError(f"unexpected protobuf wire type {wire}")
        yield field, wire, val


def _float_tensor(b: memoryview):
    """-> (name, ndarray) for FLOAT tensors, (name, None) otherwise."""
    dims: List[int] = []
    dtype = 0
    name = ""
    raw = None
    floats = None
    for field, wire, val in _walk(b):
        if field == _T_DIMS:
            if wire == 0:
                dims.append(val)
            else:
                j = 0
                while j < len(val):
                    d, j = _read_varint(val, j)
                    dims.append(d)
        elif field == _T_DTYPE:
            dtype = val
        elif field == _T_NAME:
            name = bytes(val).decode()
        elif field == _T_RAW:
            raw = val
        elif field == _T_FLOATS and wire == 2:
            floats = val
    if dtype != _FLOAT:
        return name, None
    data = raw if raw is not None else floats
    if data is None:
        return name, None
    arr = np.frombuffer(bytes(data), dtype="<f4").astype(np.float32)
    return name, arr.reshape(dims) if dims else arr


def _collect_graph(b: memoryview, out: Dict[str, np.ndarray]) -> None:
    for field, _, val in _walk(b):
        if field == _GRAPH_INIT:
            name, arr = _float_tensor(val)
            if arr is not None and name:
                out[name] = arr
        elif field == _GRAPH_NODE:
            outputs: List[str] = []
            op = ""
            attrs: List[memoryview] = []
            for f2, _, v2 in _walk(val):
                if f2 == _NODE_OUTPUT:
                    outputs.append(bytes(v2).decode())
                elif f2 == _NODE_OPTYPE:
                    op = bytes(v2).decode()
                elif f2 == _NODE_ATTR:
                    attrs.append(v2)
            for a in attrs:
                for f3, _, v3 in _walk(a):
                    if f3 == _ATTR_T and op == "Constant" and outputs:
                        _, arr = _float_tensor(v3)
                        if arr is not None:
                            out[outputs[0]] = arr
                    elif f3 in (_ATTR_G, _ATTR_GRAPHS):
                        _collect_graph(v3, out)


def read_float_tensors(path: Union[str, Path]) -> Dict[str, np.ndarray]:
    data = memoryview(Path(path).read_bytes())
    out: Dict[str, np.ndarray] = {}
    for field, _, val in _walk(data):
        if field == _MODEL_GRAPH:
            _collect_graph(val, out)
    if not out:
        raise ValueError(f"no float tensors found in {path}")
    return out


# canonical order of the v5 / 16 kHz branch (names after `If_0_then_branch__Inline_0__`)
_V5_PREFIX_16K = "If_0_then_branch__Inline_0__"
_V5_LAYOUT = (
    ("stft.forward_basis_buffer", 258 * 256),
    ("encoder.0.reparam_conv.weight", 128 * 129 * 3),
    ("encoder.0.reparam_conv.bias", 128),
    ("encoder.1.reparam_conv.weight", 64 * 128 * 3),
    ("encoder.1.reparam_conv.bias", 64),
    ("encoder.2.reparam_conv.weight", 64 * 64 * 3),
    ("encoder.2.reparam_conv.bias", 64),
    ("encoder.3.reparam_conv.weight", 128 * 64 * 3),
    ("encoder.3.reparam_conv.bias", 128),
    ("decoder.rnn.weight_ih", 512 * 128),
    ("decoder.rnn.weight_hh", 512 * 128),
    ("decoder.rnn.bias_ih", 512),
    ("decoder.rnn.bias_hh", 512),
    ("decoder.decoder.2.weight", 128),
    ("decoder.decoder.2.bias", 1),
)
V5_WEIGHT_FLOATS = 309633


def canonical_blob_v5(path: Union[str, Path]) -> np.ndarray:
    tensors = read_float_tensors(path)
    parts = []
    for name, size in _V5_LAYOUT:
        key = _V5_PREFIX_16K + name
        if key not in tensors:
            raise KeyError(f"{path}: tensor {key!r} not found (is this silero_vad_v5.onnx?)")
        t = np.ascontiguousarray(tensors[key], dtype=np.float32).reshape(-1)
        if t.size != size:
            raise ValueError(f"{path}: tensor {name} has {t.size} values, expected {size}")
        parts.append(t)
    blob = np.concatenate(parts)
    assert blob.size == V5_WEIGHT_FLOATS
    return blob


# canonical order of the v4 / 16 kHz branch.  Conv weights are top-level initializers
# (`model.*` plus the anonymous 1110..1120); the two LSTM layers are anonymous initializers
# inside the nested If subgraphs of the 16 kHz branch (W/R/B = 343/345/347 and 415/417/419).
_V4_LAYOUT = (
    ("model.feature_extractor.forward_basis_buffer", 258 * 256), ("model.adaptive_normalization.filter_", 7),
    ("model.first_layer.0.dw_conv.0.weight", 258 * 5), ("model.first_layer.0.dw_conv.0.bias", 258),
    ("model.first_layer.0.pw_conv.0.weight", 16 * 258), ("model.first_layer.0.pw_conv.0.bias", 16),
    ("model.first_layer.0.proj.weight", 16 * 258), ("model.first_layer.0.proj.bias", 16),
    ("1110", 256), ("1111", 16),
    ("model.encoder.3.0.dw_conv.0.weight", 80), ("model.encoder.3.0.dw_conv.0.bias", 16),
    ("model.encoder.3.0.pw_conv.0.weight", 512), ("model.encoder.3.0.pw_conv.0.bias", 32),
    ("model.encoder.3.0.proj.weight", 512), ("model.encoder.3.0.proj.bias", 32),
    ("1113", 1024), ("1114", 32),
    ("model.encoder.7.0.dw_conv.0.weight", 160), ("model.encoder.7.0.dw_conv.0.bias", 32),
    ("model.encoder.7.0.pw_conv.0.weight", 1024), ("model.encoder.7.0.pw_conv.0.bias", 32),
    ("1116", 1024), ("1117", 32),
    ("model.encoder.11.0.dw_conv.0.weight", 160), ("model.encoder.11.0.dw_conv.0.bias", 32),
    ("model.encoder.11.0.pw_conv.0.weight", 2048), ("model.encoder.11.0.pw_conv.0.bias", 64),
    ("model.encoder.11.0.proj.weight", 2048), ("model.encoder.11.0.proj.bias", 64),
    ("1119", 4096), ("1120", 64),
    ("343", 16384), ("345", 16384), ("347", 512), ("415", 16384), ("417", 16384), ("419", 512),
    ("model.decoder.decoder.1.weight", 64), ("model.decoder.decoder.1.bias", 1),
)
V4_WEIGHT_FLOATS = 155908


# the 8 kHz sub-model (`model_8k.*`, the graph's else-branch: any sr != 16000): same tensor shapes, its own
# anonymous 1x1 convolutions (1122..1132) and LSTM initializers (833/835/837, 905/907/909)
_V4_8K_RENAME = {"1110": "1122", "1111": "1123", "1113": "1125", "1114": "1126", "1116": "1128", "1117": "1129",
                 "1119": "1131", "1120": "1132", "343": "833", "345": "835", "347": "837", "415": "905", "417": "907",
                 "419": "909"}


def canonical_blob_v4(path: Union[str, Path], branch: str = "16k") -> np.ndarray:
    if branch not in ("16k", "8k"):
        raise ValueError(f"unknown v4 branch {branch!r}")
    tensors = read_float_tensors(path)
    parts = []
    for name, size in _V4_LAYOUT:
        if branch == "8k":
            name = _V4_8K_RENAME.get(name, name.replace("model.", "model_8k.", 1))
        if name not in tensors:
            raise KeyError(f"{path}: tensor {name!r} not found (is this silero_vad.onnx, the v4 model?)")
        t = np.ascontiguousarray(tensors[name], dtype=np.float32).reshape(-1)
        if t.size != size:
            raise ValueError(f"{path}: tensor {name} has {t.size} values, expected {size}")
        parts.append(t)
    blob = np.concatenate(parts)
    assert blob.size == V4_WEIGHT_FLOATS
    return blob
