"""ORACLE (test infrastructure, never shipped): an `onnxruntime`-shaped module.

Exposes exactly the names the reference binds at
/root/reference/src/real_time_vad/core/silero_model.py:13,315-325,345,365-366,433,559
(`InferenceSession`, `SessionOptions`, `GraphOptimizationLevel.ORT_ENABLE_ALL`,
`get_available_providers`) and executes the model with the numpy op-by-op
interpreter in oracle/onnx_interp.py.  Put `oracle/ort_shim` and `oracle` on
PYTHONPATH and the reference's UNMODIFIED Python (from /root/reference/src) runs
on top of it; that pair is the CPU oracle for probabilities and event indices.
It is NOT onnxruntime: the interpreter behind it is pinned against OpenCV's DNN module and PyTorch instead.
"""
import os
import sys

import numpy as np

_here = os.path.dirname(os.path.abspath(__file__))
_oracle_dir = os.path.dirname(os.path.dirname(_here))
if _oracle_dir not in sys.path:
    sys.path.insert(0, _oracle_dir)

from onnx_interp import OnnxInterpreter  # noqa: E402

__version__ = "0.0-oracle"


class GraphOptimizationLevel:
    ORT_DISABLE_ALL = 0
    ORT_ENABLE_BASIC = 1
    ORT_ENABLE_EXTENDED = 2
    ORT_ENABLE_ALL = 99


class SessionOptions:
    def __init__(self):
        self.inter_op_num_threads = 0
        self.intra_op_num_threads = 0
        self.graph_optimization_level = GraphOptimizationLevel.ORT_ENABLE_ALL


class _Arg:
    def __init__(self, name):
        self.name = name
        self.shape = None
        self.type = None


def get_available_providers():
    return ["CPUExecutionProvider"]


def get_device():
    return "CPU"


class InferenceSession:
    def __init__(self, path_or_bytes, sess_options=None, providers=None, **kw):
        dtype = np.float64 if os.environ.get("ORACLE_FP64") == "1" else np.float32
        self._interp = OnnxInterpreter(str(path_or_bytes), dtype=dtype)
        g = self._interp.model.graph
        self._inputs = [_Arg(n) for n in g.inputs if n not in g.initializers]
        self._outputs = [_Arg(n) for n in g.outputs]
        self._providers = ["CPUExecutionProvider"]

    def get_inputs(self):
        return list(self._inputs)

    def get_outputs(self):
        return list(self._outputs)

    def get_providers(self):
        return list(self._providers)

    def run(self, output_names, input_feed, run_options=None):
        outs = self._interp.run(input_feed, output_names)
        return [np.asarray(o, dtype=np.float32) if np.asarray(o).dtype.kind == "f" else np.asarray(o)
                for o in outs]
