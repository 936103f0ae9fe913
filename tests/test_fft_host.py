"""CPU: the per-thread routines of the double-precision FFT kernels (csrc/cvad_fft.cuh), compiled for the host.

The CUDA kernels in csrc/cvad_fftk.cuh call the same __host__ __device__ functions thread by thread; tools/fft_host_check.cpp
runs them sequentially ("16 threads, barrier, 16 threads") so that the transform, the resampler's spectral stage
(scipy.signal.resample's two Nyquist rules, /root/reference/src/real_time_vad/utils/audio.py:19-55) and the STFT packing are
checked against numpy / scipy in float64 without a GPU.  The GPU tests then only have to catch indexing mistakes of
the kernels around them."""
import ctypes
import subprocess

import numpy as np
import pytest

from conftest import ORACLE, ROOT

DP = ctypes.POINTER(ctypes.c_double)
FP = ctypes.POINTER(ctypes.c_float)


@pytest.fixture(scope="module")
def chk():
    out = ORACLE / "_build"
    out.mkdir(exist_ok=True)
    so = out / "libfft_host_check.so"
    src = ROOT / "tools" / "fft_host_check.cpp"
    hdr = ROOT / "cutter-vad_b200" / "csrc" / "cvad_fft.cuh"
    if not so.exists() or so.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-x", "c++", str(src), "-o", str(so)], check=True)
    return ctypes.CDLL(str(so))


def test_complex_256_transform_both_directions(chk):
    rng = np.random.default_rng(0)
    z = rng.standard_normal(256) + 1j * rng.standard_normal(256)
    inp = np.ascontiguousarray(np.stack([z.real, z.imag], 1).ravel())
    out = np.zeros(512)
    chk.fftchk_fft256(inp.ctypes.data_as(DP), out.ctypes.data_as(DP), 0)
    assert np.abs(out[0::2] + 1j * out[1::2] - np.fft.fft(z)).max() < 1e-12
    chk.fftchk_fft256(inp.ctypes.data_as(DP), out.ctypes.data_as(DP), 1)
    assert np.abs(out[0::2] + 1j * out[1::2] - np.fft.ifft(z) * 256).max() < 1e-12


@pytest.mark.parametrize("R", [1, 3, 6])
def test_resampler_equals_scipy_float64_rounded_once(chk, R):
    """8 / 24 / 48 kHz chunk -> 512 samples: bit-identical to scipy.signal.resample evaluated in float64 and cast to
    float32, including impulses at every phase (which exercise the two Nyquist rules)."""
    from scipy import signal
    rng = np.random.default_rng(R)
    n_in = 256 * R
    cases = [(0.3 * rng.standard_normal(n_in)).astype(np.float32) for _ in range(6)]
    for m in (0, 1, n_in // 2, n_in - 1):
        e = np.zeros(n_in, np.float32)
        e[m] = 1.0
        cases.append(e)
    cases.append(np.cos(np.pi * np.arange(n_in)).astype(np.float32))       # the source's Nyquist tone
    cases.append(np.cos(np.pi * np.arange(n_in) * 512 / n_in).astype(np.float32))   # the target's Nyquist tone
    worst32 = 0.0
    for x in cases:
        y = np.zeros(512, np.float32)
        assert chk.fftchk_resample(x.ctypes.data_as(FP), R, y.ctypes.data_as(FP)) == 0
        want64 = signal.resample(x.astype(np.float64), 512)
        assert np.abs(y - want64).max() <= 6e-8 * max(1.0, np.abs(want64).max())
        # correctly rounded: within half a float32 ulp of the float64 value (+ float64 noise around exact zeros)
        half_ulp = 0.5 * np.spacing(np.abs(want64).astype(np.float32)).astype(np.float64)
        assert np.all(np.abs(y - want64) <= half_ulp * (1 + 1e-6) + 1e-15)
        worst32 = max(worst32, float(np.abs(signal.resample(x, 512) - want64).max()))
    assert worst32 > 2e-8            # scipy's own float32 FFT is NOT correctly rounded: that is what the float64 path removes


def test_stft_equals_the_exact_hann_dft_basis(chk):
    rng = np.random.default_rng(4)
    n = np.arange(256)
    hann = 0.5 - 0.5 * np.cos(2 * np.pi * n / 256)
    k = np.arange(129)[:, None]
    bre = hann * np.cos(2 * np.pi * k * n / 256)
    bim = -hann * np.sin(2 * np.pi * k * n / 256)
    xp = (0.2 * rng.standard_normal(704)).astype(np.float32)
    re = np.zeros((8, 129))
    im = np.zeros((8, 129))
    chk.fftchk_stft(xp.ctypes.data_as(FP), re.ctypes.data_as(DP), im.ctypes.data_as(DP))
    for t in range(8):
        w = xp[64 * t:64 * t + 256].astype(np.float64)
        assert np.abs(re[t] - bre @ w).max() < 1e-12 and np.abs(im[t] - bim @ w).max() < 1e-12


def test_v4_basis_is_hann_dft_up_to_float32_rounding():
    """The premise of CVAD_MATH_FFT (cvad_create checks it too): both sub-models' STFT bases are the float32 image of
    periodic Hann x DFT-256, so the tensor-core correction term is a product with a matrix of |entries| <= 7.7e-8."""
    from conftest import V4_ONNX
    from real_time_vad.engine.onnx_weights import canonical_blob_v4
    n = np.arange(256)
    hann = 0.5 - 0.5 * np.cos(2 * np.pi * n / 256)
    k = np.arange(129)[:, None]
    exact = np.concatenate([hann * np.cos(2 * np.pi * k * n / 256), -hann * np.sin(2 * np.pi * k * n / 256)])
    for branch in ("16k", "8k"):
        blob = canonical_blob_v4(V4_ONNX, branch=branch) if branch == "8k" else canonical_blob_v4(V4_ONNX)
        basis = blob[:258 * 256].reshape(258, 256).astype(np.float64)
        assert np.abs(basis - exact).max() <= 1.0e-7, branch
