"""ORACLE (test infrastructure, never shipped): re-emit one branch of the reference's .onnx as a flat, static ONNX
model that third-party importers without control-flow support can load.

The reference's graphs are `Equal(sr, 16000) -> If(...)` with further `If` nodes and shape arithmetic
(Shape / Gather / Concat / Cast feeding Pad, Slice, Reshape, ...) inside.  OpenCV's DNN
module -- the one independent ONNX executor present in this image; onnxruntime is absent -- rejects `If`.  This module

  1. runs the op-by-op interpreter (onnx_interp.py) once on a sample feed and records the nodes it actually executed,
     in order, with the taken branch of every `If` inlined;
  2. folds everything that does not depend on the DATA of the runtime inputs (constants, and shape arithmetic: the
     output of Shape / Size is static for a fixed input shape) into initializers;
  3. serialises what is left -- the file's own compute nodes, attributes and weight tensors, untouched -- as a plain
     ModelProto (hand-rolled protobuf writer; field numbers of the public onnx.proto3).

The result computes exactly what the original computes for inputs of the sample's shape (here: 512 samples, sr = 16000).
"""
from __future__ import annotations

import struct
from typing import Any, Dict, List, Sequence, Tuple

import numpy as np

from onnx_interp import OnnxInterpreter, _Scope
from onnx_wire import Graph, Node

_SHAPE_OPS = ("Shape", "Size")


class _Tracer(OnnxInterpreter):
    def __init__(self, model_path: str):
        super().__init__(model_path, np.float32)
        self.trace: List[Tuple[Node, List[str], List[str]]] = []
        self.values: Dict[str, np.ndarray] = {}
        self.alias: Dict[str, str] = {}

    def _exec_graph(self, g: Graph, scope: _Scope) -> None:
        for node in g.nodes:
            ins = [scope.get(n) if n else None for n in node.inputs]
            if node.op == "If":
                cond = bool(np.asarray(ins[0]).reshape(-1)[0])
                branch: Graph = node.attrs["then_branch" if cond else "else_branch"]
                sub = _Scope(scope)
                for k, v in branch.initializers.items():
                    sub.vals[k] = self._f(v)
                    self.values[k] = np.asarray(sub.vals[k])
                self._exec_graph(branch, sub)
                for out_name, inner in zip(node.outputs, branch.outputs):
                    val = sub.get(inner)
                    scope.vals[out_name] = val
                    self.values[out_name] = np.asarray(val)
                    self.alias[out_name] = self.alias.get(inner, inner)      # the If's output IS the branch's value
                continue
            fn = getattr(self, "_op_" + node.op)
            res = fn(node, ins, scope)
            if not isinstance(res, (list, tuple)):
                res = [res]
            for name, val in zip(node.outputs, res):
                if name:
                    scope.vals[name] = val
                    self.values[name] = np.asarray(val)
            self.trace.append((node, [self.alias.get(i, i) for i in node.inputs], list(node.outputs)))


# ---------------------------------------------------------------- protobuf writer
def _varint(v: int) -> bytes:
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
    return _varint((fno << 3) | wt)


def _ld(fno: int, payload: bytes) -> bytes:
    return _key(fno, 2) + _varint(len(payload)) + payload


def _vi(fno: int, v: int) -> bytes:
    return _key(fno, 0) + _varint(int(v))


_DT = {np.dtype(np.float32): 1, np.dtype(np.uint8): 2, np.dtype(np.int8): 3, np.dtype(np.int32): 6, np.dtype(np.int64): 7,
       np.dtype(np.bool_): 9, np.dtype(np.float64): 11}


def _tensor(name: str, arr: np.ndarray) -> bytes:
    arr = np.ascontiguousarray(arr)
    out = b"".join(_vi(1, d) for d in arr.shape) + _vi(2, _DT[arr.dtype])
    if name:
        out += _ld(8, name.encode())
    return out + _ld(9, arr.tobytes())


def _attr(name: str, val: Any) -> bytes:
    out = _ld(1, name.encode())
    if isinstance(val, float):
        return out + _key(2, 5) + struct.pack("<f", val) + _vi(20, 1)
    if isinstance(val, (int, np.integer)):
        return out + _vi(3, int(val)) + _vi(20, 2)
    if isinstance(val, (bytes, str)):
        return out + _ld(4, val if isinstance(val, bytes) else val.encode()) + _vi(20, 3)
    if isinstance(val, np.ndarray):
        return out + _ld(5, _tensor("", val)) + _vi(20, 4)
    if isinstance(val, (list, tuple)):
        if val and isinstance(val[0], float):
            return out + b"".join(_key(7, 5) + struct.pack("<f", v) for v in val) + _vi(20, 6)
        return out + b"".join(_vi(8, int(v)) for v in val) + _vi(20, 7)
    raise TypeError(f"attribute {name}: {type(val)}")


def _node(op: str, name: str, ins: Sequence[str], outs: Sequence[str], attrs: Dict[str, Any]) -> bytes:
    out = b"".join(_ld(1, i.encode()) for i in ins) + b"".join(_ld(2, o.encode()) for o in outs)
    out += _ld(3, name.encode()) + _ld(4, op.encode())
    return out + b"".join(_ld(5, _attr(k, v)) for k, v in attrs.items())


def _value_info(name: str, arr: np.ndarray) -> bytes:
    shape = b"".join(_ld(1, _vi(1, d)) for d in arr.shape)
    tensor_type = _vi(1, _DT[arr.dtype]) + _ld(2, shape)
    return _ld(1, name.encode()) + _ld(2, _ld(1, tensor_type))


def flatten(model_path: str, feeds: Dict[str, np.ndarray], data_inputs: Sequence[str], outputs: Sequence[str] = (),
            pad_rank4: bool = False, neg_slice_as_gather: bool = False):
    """-> (serialised static ModelProto, names of the kept nodes' ops, interpreter outputs for `feeds`)."""
    tr = _Tracer(model_path)
    outs = tr.run(feeds)
    out_names = list(outputs) if outputs else list(tr.model.graph.outputs)
    tainted = set(data_inputs)
    kept: List[Tuple[Node, List[str], List[str]]] = []
    for node, ins, outs_n in tr.trace:
        dep = any(i in tainted for i in ins if i) and node.op not in _SHAPE_OPS
        if dep:
            tainted.update(o for o in outs_n if o)
            kept.append((node, ins, outs_n))
    inits: Dict[str, np.ndarray] = {}
    body = b""
    for node, ins, outs_n in kept:
        if node.op == "Pad" and pad_rank4 and tr.values[outs_n[0]].ndim < 4:
            # OpenCV's reflect padding wants NCHW: the same Pad on the tensor viewed as [1, .., dims] (an exact rewrite)
            x_shape = list(tr.values[outs_n[0]].shape)
            r = len(x_shape)
            in_shape = list(np.asarray(tr.values[ins[0]] if ins[0] in tr.values else feeds[ins[0]]).shape)
            pads = [int(v) for v in np.asarray(tr.values[ins[1]]).reshape(-1)]
            pads4 = [0] * (4 - r) + pads[:r] + [0] * (4 - r) + pads[r:]
            base = outs_n[0]
            inits[base + "/shape4"] = np.array([1] * (4 - r) + in_shape, np.int64)
            inits[base + "/pads4"] = np.array(pads4, np.int64)
            inits[base + "/shape_out"] = np.array(x_shape, np.int64)
            body += _ld(1, _node("Reshape", base + "/to4", [ins[0], base + "/shape4"], [base + "/x4"], {}))
            attrs = {k: v for k, v in node.attrs.items() if not isinstance(v, Graph)}
            body += _ld(1, _node("Pad", base + "/pad4", [base + "/x4", base + "/pads4"], [base + "/y4"], attrs))
            body += _ld(1, _node("Reshape", base + "/from4", [base + "/y4", base + "/shape_out"], [base], {}))
            continue
        if node.op == "Slice" and neg_slice_as_gather and len(ins) >= 5 and ins[4] and ins[4] not in tainted:
            steps = [int(v) for v in np.asarray(tr.values[ins[4]]).reshape(-1)]
            if len(steps) == 1 and steps[0] < 0:
                # reversed slice (the v4 graph builds its reflect padding from these): the same elements by index
                axis = int(np.asarray(tr.values[ins[3]]).reshape(-1)[0])
                x = np.asarray(tr.values[ins[0]] if ins[0] in tr.values else feeds[ins[0]])
                n = x.shape[axis]
                start = int(np.asarray(tr.values[ins[1]]).reshape(-1)[0])
                end = int(np.asarray(tr.values[ins[2]]).reshape(-1)[0])
                idx = np.arange(n)[slice(start if start > -(1 << 62) else None, end if end > -(1 << 62) else None, steps[0])]
                assert np.array_equal(np.take(x, idx, axis=axis), tr.values[outs_n[0]])
                inits[outs_n[0] + "/idx"] = idx.astype(np.int64)
                body += _ld(1, _node("Gather", outs_n[0] + "/gather", [ins[0], outs_n[0] + "/idx"], [outs_n[0]], {"axis": axis}))
                continue
        for i in ins:
            if i and i not in tainted and i not in inits:
                v = tr.values[i] if i in tr.values else (
                    tr.model.graph.initializers[i] if i in tr.model.graph.initializers else np.asarray(feeds[i]))
                inits[i] = np.asarray(v)
        attrs = {k: v for k, v in node.attrs.items() if not isinstance(v, Graph)}
        body += _ld(1, _node(node.op, node.name or node.outputs[0], ins, outs_n, attrs))
    for k in out_names:                       # graph outputs that are If outputs: name the branch's value
        src = tr.alias.get(k, k)
        if src != k:
            body += _ld(1, _node("Identity", f"output/{k}", [src], [k], {}))
    graph = body + _ld(2, b"flattened")
    graph += b"".join(_ld(5, _tensor(k, v)) for k, v in inits.items())
    graph += b"".join(_ld(11, _value_info(k, np.asarray(feeds[k]))) for k in data_inputs)
    graph += b"".join(_ld(12, _value_info(k, tr.values[k])) for k in out_names)
    model = _vi(1, 8) + _ld(2, b"cutter-vad-b200-oracle") + _ld(8, _vi(2, tr.model.opset or 16)) + _ld(7, graph)
    return model, [n.op for n, _, _ in kept], dict(zip(tr.model.graph.outputs, outs))


class OpenCVSession:
    """The flattened graph inside OpenCV's DNN module: an `InferenceSession`-like `run` from a third-party ONNX executor.
    Two exact rewrites adapt the file's nodes to OpenCV's importer: the reflect Pad runs on a rank-4 view of its input,
    and reversed Slices (step -1) become Gathers of the same elements; everything else -- Conv, LSTM, Pow, Sqrt, Log,
    ReduceMean, Sigmoid, ... with the file's attributes and weight tensors -- is executed as the file states it."""

    def __init__(self, model_path: str, feeds: Dict[str, np.ndarray], data_inputs: Sequence[str]):
        import cv2
        self.data_inputs = list(data_inputs)
        blob, self.ops, self.sample_outputs = flatten(model_path, feeds, data_inputs, pad_rank4=True, neg_slice_as_gather=True)
        self.output_names = list(self.sample_outputs.keys())
        self.shapes = {k: np.asarray(v).shape for k, v in self.sample_outputs.items()}
        self.net = cv2.dnn.readNetFromONNX(np.frombuffer(blob, np.uint8))
        self.net.setPreferableBackend(cv2.dnn.DNN_BACKEND_OPENCV)
        self.net.setPreferableTarget(cv2.dnn.DNN_TARGET_CPU)

    def run(self, feeds: Dict[str, np.ndarray]) -> List[np.ndarray]:
        for k in self.data_inputs:
            self.net.setInput(np.ascontiguousarray(feeds[k], dtype=np.float32), k)
        outs = self.net.forward(self.output_names)
        return [np.asarray(o, np.float32).reshape(self.shapes[n]).copy() for n, o in zip(self.output_names, outs)]
