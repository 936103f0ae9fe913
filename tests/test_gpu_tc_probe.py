"""GPU: tcgen05 hardware probe -- one BF16 tensor-core GEMM (SW128 K-major smem descriptors, TMEM
accumulator, tcgen05.ld epilogue) against numpy.  Guards the descriptor encodings in csrc/cvad_tc.cuh."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bf16_bits(x: np.ndarray) -> np.ndarray:
    """float32 -> bf16 (round to nearest even) as uint16 bit patterns."""
    u = x.astype(np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + 0x7FFF
    return ((u + r) >> 16).astype(np.uint16)


def _bf16_val(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << 16).view(np.float32)


def test_tcgen05_probe_matches_numpy():
    from real_time_vad.engine import capi
    L = capi.dev_lib()
    rng = np.random.default_rng(0)
    A = rng.standard_normal((128, 256)).astype(np.float32)
    B = rng.standard_normal((32, 256)).astype(np.float32)
    a, b = _bf16_bits(A), _bf16_bits(B)
    D = np.zeros((128, 32), np.float32)
    rc = L.cvad_tc_probe(0, a.ctypes.data, b.ctypes.data, D.ctypes.data)
    assert rc == 0, L.cvad_dev_last_error()
    assert D[0, 0] != -12345.0, "tensor-core MMA never completed (descriptor problem)"
    want = _bf16_val(a).astype(np.float64) @ _bf16_val(b).astype(np.float64).T
    err = np.abs(D - want).max()
    assert err < 2e-4 * np.abs(want).max(), f"max err {err}"


def test_tcgen05_mn_major_b_operand_matches_numpy():
    """B operand stored MN-major SWIZZLE_64B (N contiguous: an epilogue thread that owns one channel x 8 streams stores one
    16-byte chunk), atoms 512 B apart along N and 1,536 B along K; second product through a descriptor that starts one N
    atom further (a conv tap).  Guards tc::mn64_offset / smem_desc_mn64 / kIdescBMajorMN in csrc/cvad_tc.cuh."""
    from real_time_vad.engine import capi
    L = capi.dev_lib()
    rng = np.random.default_rng(1)
    A = rng.standard_normal((128, 64)).astype(np.float32)
    B = rng.standard_normal((96, 64)).astype(np.float32)
    a, b = _bf16_bits(A), _bf16_bits(B)
    D = np.zeros((128, 160), np.float32)
    rc = L.cvad_tc_probe_mn(0, a.ctypes.data, b.ctypes.data, D.ctypes.data)
    assert rc == 0, L.cvad_dev_last_error()
    assert D[0, 0] != -12345.0, "tensor-core MMA never completed (descriptor problem)"
    want = _bf16_val(a).astype(np.float64) @ _bf16_val(b).astype(np.float64).T
    assert np.abs(D[:, :96] - want).max() < 2e-4 * np.abs(want).max()
    assert np.abs(D[:, 96:] - want[:, 32:]).max() < 2e-4 * np.abs(want).max()
