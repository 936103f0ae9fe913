"""ORACLE (test infrastructure, never shipped): op-by-op ONNX interpreter in numpy.

Executes the reference's own model files node by node under ONNX opset-16
semantics, so the only thing trusted is the published meaning of each op, not a
hand-written restatement of the network.  This is the stand-in for the
`onnxruntime.InferenceSession.run` call at
/root/reference/src/real_time_vad/core/silero_model.py:433 (onnxruntime itself is
an unpinned third-party wheel, `pyproject.toml:32` `onnxruntime>=1.10.0`, absent
from this image).

Pinning: the reference's tests mock every `session.run` (tests/test_silero_model.py:278-292) and hold no golden
probability vector, and onnxruntime itself cannot be run here.  This interpreter is therefore checked against a
third-party ONNX executor that can: OpenCV's DNN module executing the same graphs (oracle/onnx_flatten.py;
tests/test_oracle_pinning.py: v5 goldens <= 5e-6, 40 stateful frames <= 5e-5, both v4 branches <= 2e-5), and against
PyTorch's own conv1d / LSTMCell (oracle/torch_reference.py).  Not pinned: onnxruntime's own rounding.

`dtype=np.float32` reproduces the reference's arithmetic type;
`dtype=np.float64` is the high-precision run used to bound FP32 rounding.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from onnx_wire import Graph, Model, Node, load_model

INT64_MIN = -(1 << 63)
INT64_MAX = (1 << 63) - 1


class _Scope:
    """Value environment with outer-scope lookup (ONNX subgraphs see their parents)."""

    def __init__(self, parent: Optional["_Scope"] = None):
        self.vals: Dict[str, Any] = {}
        self.parent = parent

    def get(self, name: str):
        s = self
        while s is not None:
            if name in s.vals:
                return s.vals[name]
            s = s.parent
        raise KeyError(name)

    def has(self, name: str) -> bool:
        s = self
        while s is not None:
            if name in s.vals:
                return True
            s = s.parent
        return False


def _sigmoid(x):
    with np.errstate(over="ignore"):
        return 1.0 / (1.0 + np.exp(-x))


class OnnxInterpreter:
    def __init__(self, model_path: str, dtype=np.float32):
        self.model: Model = load_model(model_path)
        self.dtype = np.dtype(dtype)
        self._const_cache: Dict[int, Any] = {}

    # ------------------------------------------------------------ helpers
    def _f(self, arr: np.ndarray) -> np.ndarray:
        """Float tensors take the interpreter's working precision."""
        if isinstance(arr, np.ndarray) and arr.dtype.kind == "f" and arr.dtype != self.dtype:
            return arr.astype(self.dtype)
        return arr

    # ------------------------------------------------------------ run
    def run(self, feeds: Dict[str, np.ndarray], outputs: Optional[Sequence[str]] = None) -> List[np.ndarray]:
        g = self.model.graph
        scope = _Scope()
        for k, v in g.initializers.items():
            scope.vals[k] = self._f(v)
        for k, v in feeds.items():
            scope.vals[k] = self._f(np.asarray(v))
        self._exec_graph(g, scope)
        names = list(outputs) if outputs is not None else g.outputs
        return [np.asarray(scope.get(n)) for n in names]

    def _exec_graph(self, g: Graph, scope: _Scope) -> None:
        for node in g.nodes:
            fn = getattr(self, "_op_" + node.op, None)
            if fn is None:
                raise NotImplementedError(f"ONNX op {node.op} ({node.name})")
            ins = [scope.get(n) if n else None for n in node.inputs]
            res = fn(node, ins, scope)
            if not isinstance(res, (list, tuple)):
                res = [res]
            for name, val in zip(node.outputs, res):
                if name:
                    scope.vals[name] = val

    # ------------------------------------------------------------ control / shape ops
    def _op_Constant(self, node: Node, ins, scope):
        key = id(node)
        if key not in self._const_cache:
            a = node.attrs
            if "value" in a:
                v = self._f(np.asarray(a["value"]))
            elif "value_float" in a:
                v = np.asarray(a["value_float"], dtype=self.dtype)
            elif "value_int" in a:
                v = np.asarray(a["value_int"], dtype=np.int64)
            elif "value_ints" in a:
                v = np.asarray(a["value_ints"], dtype=np.int64)
            elif "value_floats" in a:
                v = np.asarray(a["value_floats"], dtype=self.dtype)
            else:
                raise NotImplementedError(f"Constant attrs {list(a)}")
            self._const_cache[key] = v
        return self._const_cache[key]

    def _op_ConstantOfShape(self, node, ins, scope):
        shape = [int(x) for x in np.asarray(ins[0]).reshape(-1)]
        val = node.attrs.get("value")
        if val is None:
            return np.zeros(shape, dtype=self.dtype)
        val = self._f(np.asarray(val))
        return np.full(shape, val.reshape(-1)[0], dtype=val.dtype)

    def _op_Identity(self, node, ins, scope):
        return ins[0]

    def _op_Cast(self, node, ins, scope):
        to = node.attrs["to"]
        table = {1: self.dtype, 6: np.int32, 7: np.int64, 9: np.bool_, 11: np.float64, 10: np.float16}
        return np.asarray(ins[0]).astype(table[to])

    def _op_Shape(self, node, ins, scope):
        return np.asarray(np.asarray(ins[0]).shape, dtype=np.int64)

    def _op_Size(self, node, ins, scope):
        return np.asarray(np.asarray(ins[0]).size, dtype=np.int64)

    def _op_Gather(self, node, ins, scope):
        axis = node.attrs.get("axis", 0)
        return np.take(np.asarray(ins[0]), np.asarray(ins[1]).astype(np.int64), axis=axis)

    def _op_Equal(self, node, ins, scope):
        return np.equal(ins[0], ins[1])

    def _op_Not(self, node, ins, scope):
        return np.logical_not(ins[0])

    def _op_If(self, node, ins, scope):
        cond = bool(np.asarray(ins[0]).reshape(-1)[0])
        branch: Graph = node.attrs["then_branch" if cond else "else_branch"]
        sub = _Scope(scope)
        for k, v in branch.initializers.items():
            sub.vals[k] = self._f(v)
        self._exec_graph(branch, sub)
        return [sub.get(n) for n in branch.outputs]

    def _op_Reshape(self, node, ins, scope):
        x = np.asarray(ins[0])
        shape = [int(s) for s in np.asarray(ins[1]).reshape(-1)]
        if not node.attrs.get("allowzero", 0):
            shape = [x.shape[i] if s == 0 else s for i, s in enumerate(shape)]
        return x.reshape(shape)

    def _op_Unsqueeze(self, node, ins, scope):
        x = np.asarray(ins[0])
        axes = [int(a) for a in np.asarray(ins[1]).reshape(-1)]
        out_rank = x.ndim + len(axes)
        axes = sorted(a + out_rank if a < 0 else a for a in axes)
        for a in axes:
            x = np.expand_dims(x, a)
        return x

    def _op_Squeeze(self, node, ins, scope):
        x = np.asarray(ins[0])
        if len(ins) > 1 and ins[1] is not None:
            axes = tuple(int(a) for a in np.asarray(ins[1]).reshape(-1))
            return np.squeeze(x, axis=axes)
        return np.squeeze(x)

    def _op_Concat(self, node, ins, scope):
        return np.concatenate([np.asarray(i) for i in ins], axis=node.attrs["axis"])

    def _op_Slice(self, node, ins, scope):
        x = np.asarray(ins[0])
        starts = [int(v) for v in np.asarray(ins[1]).reshape(-1)]
        ends = [int(v) for v in np.asarray(ins[2]).reshape(-1)]
        axes = ([int(v) for v in np.asarray(ins[3]).reshape(-1)]
                if len(ins) > 3 and ins[3] is not None else list(range(len(starts))))
        steps = ([int(v) for v in np.asarray(ins[4]).reshape(-1)]
                 if len(ins) > 4 and ins[4] is not None else [1] * len(starts))
        sl = [slice(None)] * x.ndim
        for st, en, ax, sp in zip(starts, ends, axes, steps):
            ax = ax + x.ndim if ax < 0 else ax
            dim = x.shape[ax]
            # ONNX clamps; python slices clamp the same way except for the
            # "run to the beginning with negative step" sentinel (INT64_MIN).
            if sp < 0:
                st_c = min(max(st + dim if st < 0 else st, -1), dim - 1)
                if en <= -dim - 1 or en == INT64_MIN:
                    en_c = None
                else:
                    en_c = min(max(en + dim if en < 0 else en, -1), dim - 1)
                    if en_c == -1:
                        en_c = None
                sl[ax] = slice(st_c, en_c, sp)
            else:
                st_c = min(max(st + dim if st < 0 else st, 0), dim)
                en_c = min(max(en + dim if en < 0 else en, 0), dim)
                sl[ax] = slice(st_c, en_c, sp)
        return x[tuple(sl)]

    def _op_Transpose(self, node, ins, scope):
        perm = node.attrs.get("perm")
        return np.transpose(np.asarray(ins[0]), perm)

    def _op_Pad(self, node, ins, scope):
        x = np.asarray(ins[0])
        pads = [int(v) for v in np.asarray(ins[1]).reshape(-1)]
        mode = node.attrs.get("mode", b"constant")
        mode = mode.decode() if isinstance(mode, bytes) else mode
        n = x.ndim
        width = [(pads[i], pads[i + n]) for i in range(n)]
        if mode == "constant":
            cval = 0.0
            if len(ins) > 2 and ins[2] is not None:
                cval = np.asarray(ins[2]).reshape(-1)[0]
            return np.pad(x, width, mode="constant", constant_values=cval)
        if mode == "reflect":
            return np.pad(x, width, mode="reflect")
        if mode == "edge":
            return np.pad(x, width, mode="edge")
        raise NotImplementedError(mode)

    # ------------------------------------------------------------ arithmetic
    def _op_Pow(self, node, ins, scope):
        y = np.asarray(ins[1])
        x = np.asarray(ins[0])
        if y.size == 1 and float(y.reshape(-1)[0]) == 2.0:
            return x * x
        return np.power(x, y.astype(x.dtype))

    def _op_Add(self, node, ins, scope):
        return np.add(ins[0], ins[1])

    def _op_Mul(self, node, ins, scope):
        return np.multiply(ins[0], ins[1])

    def _op_Neg(self, node, ins, scope):
        return np.negative(ins[0])

    def _op_Log(self, node, ins, scope):
        return np.log(ins[0])

    def _op_Sqrt(self, node, ins, scope):
        return np.sqrt(ins[0])

    def _op_Relu(self, node, ins, scope):
        return np.maximum(ins[0], 0)

    def _op_Sigmoid(self, node, ins, scope):
        x = np.asarray(ins[0])
        return _sigmoid(x).astype(x.dtype)

    def _op_ReduceMean(self, node, ins, scope):
        x = np.asarray(ins[0])
        axes = node.attrs.get("axes")
        keep = bool(node.attrs.get("keepdims", 1))
        axes = tuple(axes) if axes is not None else None
        return np.mean(x, axis=axes, keepdims=keep, dtype=x.dtype)

    def _op_Conv(self, node, ins, scope):
        x = np.asarray(ins[0])
        w = np.asarray(ins[1])
        b = np.asarray(ins[2]) if len(ins) > 2 and ins[2] is not None else None
        a = node.attrs
        group = a.get("group", 1)
        if x.ndim != 3:
            raise NotImplementedError(f"Conv rank {x.ndim}")
        stride = a.get("strides", [1])[0]
        dil = a.get("dilations", [1])[0]
        pads = a.get("pads", [0, 0])
        k = w.shape[2]
        if pads[0] or pads[1]:
            x = np.pad(x, ((0, 0), (0, 0), (pads[0], pads[1])))
        N, C, L = x.shape
        M = w.shape[0]
        span = (k - 1) * dil + 1
        lout = (L - span) // stride + 1
        # windows [N, C, lout, k]
        idx = (np.arange(lout)[:, None] * stride) + (np.arange(k)[None, :] * dil)
        win = x[:, :, idx]
        cg = C // group
        mg = M // group
        if group == 1:
            out = np.einsum("nclk,mck->nml", win, w, optimize=True)
        elif cg == 1 and mg == 1:  # depthwise
            out = np.einsum("nclk,ck->ncl", win, w[:, 0, :], optimize=True)
        else:
            outs = []
            for gi in range(group):
                outs.append(np.einsum("nclk,mck->nml", win[:, gi * cg:(gi + 1) * cg],
                                      w[gi * mg:(gi + 1) * mg], optimize=True))
            out = np.concatenate(outs, axis=1)
        out = out.astype(x.dtype, copy=False)
        if b is not None:
            out = out + b[None, :, None]
        return out

    def _op_LSTM(self, node, ins, scope):
        a = node.attrs
        direction = a.get("direction", b"forward")
        direction = direction.decode() if isinstance(direction, bytes) else direction
        if direction != "forward":
            raise NotImplementedError(direction)
        if a.get("layout", 0) != 0:
            raise NotImplementedError("LSTM layout=1")
        X = np.asarray(ins[0])
        W = np.asarray(ins[1])[0]
        R = np.asarray(ins[2])[0]
        H = a["hidden_size"]
        B = np.asarray(ins[3])[0] if len(ins) > 3 and ins[3] is not None else np.zeros(8 * H, X.dtype)
        T, Bn, _ = X.shape
        h = np.asarray(ins[5])[0] if len(ins) > 5 and ins[5] is not None else np.zeros((Bn, H), X.dtype)
        c = np.asarray(ins[6])[0] if len(ins) > 6 and ins[6] is not None else np.zeros((Bn, H), X.dtype)
        Wb, Rb = B[:4 * H], B[4 * H:]
        ys = []
        for t in range(T):
            gates = X[t] @ W.T + h @ R.T + Wb + Rb  # iofc
            i = _sigmoid(gates[:, 0:H])
            o = _sigmoid(gates[:, H:2 * H])
            f = _sigmoid(gates[:, 2 * H:3 * H])
            g = np.tanh(gates[:, 3 * H:4 * H])
            c = f * c + i * g
            h = o * np.tanh(c)
            ys.append(h)
        Y = np.stack(ys, 0)[:, None].astype(X.dtype)
        return [Y, h[None].astype(X.dtype), c[None].astype(X.dtype)]
