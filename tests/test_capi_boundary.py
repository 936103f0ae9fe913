"""CPU: the C-ABI shared library loads, exports every symbol include/*.h declares, has the
struct layouts the ctypes binding assumes, and fails loudly without a GPU (no fallback)."""
import ctypes as C
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from conftest import ROOT

HEADER = ROOT / "include" / "cutter_vad_b200.h"


DEV_HEADER = ROOT / "include" / "cutter_vad_b200_dev.h"


def _declared_functions(header=None):
    text = (header or HEADER).read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cvad_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from real_time_vad.engine import capi
    L = capi.lib()
    names = _declared_functions()
    assert len(names) >= 17
    for n in names:
        assert hasattr(L, n), f"libcvad_b200.so does not export {n}"
    assert set(names) == set(capi.EXPORTS)
    assert L.cvad_abi_version() == 4


def test_development_probes_live_in_their_own_library():
    """The tcgen05 probes (hardware probe, MMA issue rate, bulk-copy rate) are test infrastructure: declared in
    include/cutter_vad_b200_dev.h, exported by libcvad_b200_dev.so only -- the product library holds none of them."""
    from real_time_vad.engine import capi
    names = _declared_functions(DEV_HEADER)
    assert set(names) == set(capi.DEV_EXPORTS)
    D = capi.dev_lib()
    L = capi.lib()
    for n in names:
        assert hasattr(D, n), f"libcvad_b200_dev.so does not export {n}"
        assert not hasattr(L, n), f"libcvad_b200.so still exports the development hook {n}"


def test_ctypes_struct_layout_matches_the_header(tmp_path):
    from real_time_vad.engine import capi
    src = tmp_path / "layout.c"
    src.write_text(f'#include <stdio.h>\n#include <stddef.h>\n#include "{HEADER}"\n'
                   'int main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(cvad_step_args), sizeof(cvad_event),'
                   'offsetof(cvad_step_args, audio), offsetof(cvad_step_args, stream_stride),'
                   'offsetof(cvad_step_args, probs_out), offsetof(cvad_step_args, n_events_out),'
                   'offsetof(cvad_event, stream_frame));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    A, E = capi.StepArgs, capi.Event
    want = [C.sizeof(A), C.sizeof(E), A.audio.offset, A.stream_stride.offset, A.probs_out.offset,
            A.n_events_out.offset, E.stream_frame.offset]
    assert got == want


def test_header_cites_the_reference_interface():
    text = HEADER.read_text()
    for cite in ("silero_model.py:433", "vad_wrapper.py:627-644", "audio.py:164-190", "silero_model.py:790-923"):
        assert cite in text


def test_weight_blob_from_packaged_onnx():
    from real_time_vad.engine.onnx_weights import canonical_blob_v5, V5_WEIGHT_FLOATS
    from conftest import V5_ONNX
    from vad_oracle import v5_blob
    blob = canonical_blob_v5(V5_ONNX)
    assert blob.dtype == np.float32 and blob.size == V5_WEIGHT_FLOATS == 309633
    assert np.array_equal(blob, v5_blob(str(V5_ONNX)))          # independent readers agree
    assert not blob[129 * 256:130 * 256].any()                   # imaginary row of bin 0 is exactly zero
    assert not blob[257 * 256:258 * 256].any()                   # imaginary row of bin 128 too


@pytest.mark.skipif(__import__("shutil").which("nvidia-smi") is not None, reason="only meaningful without a GPU")
def test_no_cpu_fallback_engine_fails_loudly():
    from real_time_vad.engine import capi
    from real_time_vad.engine.stream_engine import EngineError, StreamEngine
    assert capi.lib().cvad_device_count() == 0
    with pytest.raises(EngineError) as ei:
        StreamEngine("v5", max_streams=4)
    assert ei.value.code == capi.E_NOGPU
    from real_time_vad import VADError, VADWrapper
    with pytest.raises(VADError, match="Failed to initialize VAD wrapper"):
        VADWrapper()


def test_missing_library_is_an_error_not_a_fallback(tmp_path, monkeypatch):
    from real_time_vad.engine import capi
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setenv("CVAD_B200_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(capi.EngineLibraryMissing):
        capi.lib()
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.delenv("CVAD_B200_LIB")
    capi.lib()


@pytest.mark.parametrize("rate,n_in", [(8000, 256), (24000, 768), (48000, 1536)])
def test_resample_operator_equals_scipy(rate, n_in):
    """The closed-form R^T the engine uploads == scipy.signal.resample applied to the identity
    (incl. both Nyquist rules), i.e. AudioUtils.resample_audio on one chunk (audio.py:19-55)."""
    from scipy import signal
    from real_time_vad.engine.stream_engine import StreamEngine
    rt = StreamEngine.resample_matrix(rate)
    assert rt.shape == (n_in, 512)
    R = signal.resample(np.eye(n_in), 512, axis=0)
    assert np.abs(rt.T - R).max() < 6e-8
    x = np.random.default_rng(rate).standard_normal(n_in).astype(np.float32)
    y = signal.resample(x, 512).astype(np.float32)
    assert np.abs(x.astype(np.float64) @ rt.astype(np.float64) - y).max() < 2e-6
