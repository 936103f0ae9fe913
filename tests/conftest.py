"""Test configuration: paths, the `gpu` marker, shared fixtures.

`-m "not gpu"` tests run on a CPU-only box (oracle vs golden vectors, host logic,
C-ABI symbol checks).  `-m gpu` tests are the parity tests proper: they drive the
CUDA engine through the C ABI and compare with the oracle.
"""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
PKG = ROOT / "cutter-vad_b200"
ORACLE = ROOT / "oracle"
GOLDEN = ROOT / "tests" / "golden"
for p in (str(PKG), str(ORACLE), str(ROOT)):
    if p not in sys.path:
        sys.path.insert(0, p)

MODELS = PKG / "real_time_vad" / "models"
V5_ONNX = MODELS / "silero_vad_v5.onnx"
V4_ONNX = MODELS / "silero_vad.onnx"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `-m gpu` on the GPU box)")


def require_cv2():
    """OpenCV's DNN module is the third-party ONNX executor the oracle is pinned against.  It ships with this image (here
    and on the GPU box); if it ever does not, the pin must FAIL, not skip -- a skipped pin leaves the oracle unpinned
    without anyone noticing."""
    try:
        import cv2  # noqa: F401
    except ImportError as e:                                   # pragma: no cover
        pytest.fail(f"cv2 (OpenCV DNN) is required to pin the oracle against a third-party ONNX executor: {e}")


def synth_streams(n_streams: int, n_samples: int, seed: int = 0) -> np.ndarray:
    """Per-stream mix of noise floors and gated harmonic 'voice' bursts (the recipe of
    examples/probability_demo.py:60-67 in the reference, seeded), so that probabilities
    cover the whole range and the state machine fires."""
    out = np.zeros((n_streams, n_samples), np.float32)
    t = np.arange(n_samples) / 16000.0
    for s in range(n_streams):
        rng = np.random.default_rng(seed * 100003 + s)
        sigma = (0.005, 0.02, 0.1)[s % 3]
        x = sigma * rng.standard_normal(n_samples)
        level = (0.3, 0.7)[(s // 3) % 2]
        f0 = rng.uniform(110, 220)
        voice = level * (0.4 * np.sin(2 * np.pi * f0 * t) + 0.3 * np.sin(2 * np.pi * 2 * f0 * t)
                         + 0.2 * np.sin(2 * np.pi * 4 * f0 * t)) * (1 + 0.5 * np.sin(2 * np.pi * 4 * t))
        gate = np.zeros(n_samples)
        pos = int(rng.uniform(0, 0.5) * 16000)
        while pos < n_samples:
            on = int(rng.uniform(0.5, 3.0) * 16000)
            gate[pos:pos + on] = 1.0
            pos += on + int(rng.uniform(0.5, 3.0) * 16000)
        out[s] = (x + gate * (voice + 0.1 * rng.standard_normal(n_samples))).astype(np.float32)
    return out


@pytest.fixture(scope="session")
def ref_lib():
    from vad_oracle import RefLib
    return RefLib()


@pytest.fixture(scope="session")
def ref_v5(ref_lib):
    from vad_oracle import RefV5, v5_blob
    return RefV5(v5_blob(str(V5_ONNX)), ref_lib)


@pytest.fixture(scope="session")
def engine_factory():
    from real_time_vad.engine.stream_engine import StreamEngine
    made = []

    def make(max_streams=64, **kw):
        e = StreamEngine(max_streams=max_streams, **kw)
        made.append(e)
        return e

    yield make
    for e in made:
        e.close()


@pytest.fixture(scope="session")
def ref_v4(ref_lib):
    from vad_oracle import RefV4, v4_blob
    return RefV4(v4_blob(str(V4_ONNX)), ref_lib)
