"""GPU: fused 8/24/48 kHz -> 16 kHz resampling.

Oracle: AudioUtils.resample_audio (= scipy.signal.resample, FFT method) applied to each
chunk of 512*sr/16000 source samples, then the 16 kHz model -- the definition SURVEY.md
section 0 fact 4 / section 7 gives, since the reference itself never resamples on its path."""
import numpy as np
import pytest

from conftest import synth_streams

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(autouse=True, params=["tc", "fp32", "tc16"])
def math_mode(request, monkeypatch):
    """Both builds: resampler + model GEMMs on the tcgen05 tensor cores (BF16 3-way split), or FP32 FMA."""
    monkeypatch.setenv("CVAD_MATH", request.param)
    return request.param


def _source_audio(n, T, rate, seed):
    """Band-limited source-rate signal: the 16 kHz recipe upsampled by scipy, plus a little wideband noise."""
    from scipy import signal
    n_in = rate * 512 // 16000
    base = synth_streams(n, 512 * T, seed=seed)
    up = signal.resample(base, n_in * T, axis=1).astype(np.float32)
    up += (0.003 * np.random.default_rng(seed).standard_normal(up.shape)).astype(np.float32)
    return up, n_in


def _oracle_16k(x, n_in):
    from vad_oracle import resample
    n, L = x.shape
    T = L // n_in
    y = np.zeros((n, 512 * T), np.float32)
    rate = n_in * 16000 // 512
    for s in range(n):
        for j in range(T):
            y[s, j * 512:(j + 1) * 512] = resample(x[s, j * n_in:(j + 1) * n_in], rate, 16000)
    return y


@pytest.mark.parametrize("rate", [8000, 24000, 48000])
def test_resampled_streams_match_scipy_then_model(engine_factory, ref_v5, rate):
    eng = engine_factory(128)
    eng.reset()
    eng.configure(enable_denoising=False)
    n, T = 37, 12
    x, n_in = _source_audio(n, T, rate, seed=rate // 1000)
    y = _oracle_16k(x, n_in)
    want, _, _ = ref_v5.run(y, T, denoise=False)
    r1 = eng.step(x[:, :n_in * 5], src_rate=rate)            # two calls: state carries across resampled steps
    r2 = eng.step(x[:, n_in * 5:], src_rate=rate)
    got = np.concatenate([r1.probs, r2.probs], axis=1)
    assert got.shape == (n, T)
    assert np.abs(got - want).max() <= TOL


def test_resampled_int16_with_noise_gate(engine_factory, ref_v5):
    """48 kHz int16 PCM, gate on.  The gate (|x| > 0.01) is discontinuous, so a resampled sample
    within 1e-6 of the threshold may fall on the other side than scipy's FP32 FFT puts it; such a
    flip moves one sample by 0.01.  Assert the 1e-4 bar on (nearly) all frames and a loose bound on all."""
    from real_time_vad.engine import capi
    eng = engine_factory(128)
    eng.reset()
    eng.configure(enable_denoising=True)
    n, T = 33, 10
    x, n_in = _source_audio(n, T, 48000, seed=5)
    q = np.clip(np.round(x * 32767.0), -32768, 32767).astype(np.int16)
    xf = q.astype(np.float32) / np.float32(32767.0)
    want, _, _ = ref_v5.run(_oracle_16k(xf, n_in), T, denoise=True)
    r = eng.step(q, src_rate=48000, pcm_format=capi.PCM_S16_32767)
    err = np.abs(r.probs - want)
    assert (err <= TOL).mean() >= 0.99
    assert err.max() <= 5e-3


def test_resampled_v4_model(engine_factory, ref_v4):
    """BASELINE config 3: v4 model behind the resampler.

    24 kHz source (down-sampling keeps the whole 0-8 kHz band): the 1e-4 bar holds.
    8 kHz source (config 3 proper): up-sampled audio has an EMPTY 4-8 kHz band, and v4 feeds
    log(1 + 2^20 |STFT|) of those ~1e-7 bins into the network, so rounding-level differences
    between two correct resamplers (scipy's FP32 FFT vs this GEMM) are amplified by ~1e4.  The same
    happens between FP32 and FP64 evaluation of the unmodified graph (7e-4 on a pure tone,
    DESIGN.md section 5), i.e. onnxruntime would not reproduce itself to 1e-4 there either.
    tests/test_oracle_pinning.py::test_v4_is_ill_conditioned_on_band_limited_input shows two CPU
    resamplers that agree to 2.4e-7 on the audio disagree by 2.4e-3 on v4's output (v5: 1e-6).
    Asserted here: median 1e-5, 1e-4 on most frames, 1e-2 worst case."""
    eng = engine_factory(128, model_version="v4")
    eng.configure(enable_denoising=False)
    n, T = 40, 9
    eng.reset()
    x, n_in = _source_audio(n, T, 24000, seed=8)
    want, _, _ = ref_v4.run(_oracle_16k(x, n_in), T, denoise=False)
    r = eng.step(x, src_rate=24000)
    assert np.abs(r.probs - want).max() <= TOL
    eng.reset()
    x, n_in = _source_audio(n, T, 8000, seed=8)
    assert n_in == 256
    want, _, _ = ref_v4.run(_oracle_16k(x, n_in), T, denoise=False)
    r = eng.step(x, src_rate=8000)
    err = np.abs(r.probs - want)
    assert np.median(err) <= 1e-5 and (err <= TOL).mean() >= 0.7 and err.max() <= 1e-2


def test_resampler_argument_checks(engine_factory):
    from real_time_vad.engine.stream_engine import EngineError
    eng = engine_factory(128)
    with pytest.raises(EngineError):
        eng.step(np.zeros((2, 4096), np.float32), src_rate=44100)
    with pytest.raises(EngineError):                            # frame_len/hop must be the chunk size
        a, keep = eng._args(np.zeros((2, 1536), np.float32), None, None, 1, 512, 512, 0, 48000)
        eng._check(eng._L.cvad_step(eng.handle, __import__("ctypes").byref(a)))


def test_batched_manager_with_24k_streams(ref_v5):
    from real_time_vad import BatchedVADManager, SampleRate, VADConfig
    n, T = 9, 14
    x, n_in = _source_audio(n, T, 24000, seed=24)
    cfg = VADConfig(sample_rate=SampleRate.SAMPLERATE_24, enable_denoising=False, vad_start_probability=0.5,
                    vad_end_probability=0.35, voice_start_frame_count=2, voice_end_frame_count=3)
    mgr = BatchedVADManager(max_streams=16, source_rate=24000)
    assert mgr.frame_len == 768
    ids = [mgr.open_stream(cfg) for _ in range(n)]
    probs = {i: [] for i in ids}
    for lo in range(0, x.shape[1], 1000):                        # arbitrary push sizes
        for k, sid in enumerate(ids):
            mgr.push(sid, x[k, lo:lo + 1000])
        out = mgr.step()
        for sid, p in out.probabilities.items():
            probs[sid].append(p)
    want, _, _ = ref_v5.run(_oracle_16k(x, n_in), T, denoise=False)
    for k, sid in enumerate(ids):
        got = np.concatenate(probs[sid])
        assert len(got) == T and np.abs(got - want[k]).max() <= TOL
    mgr.close()


def test_mixed_source_rates_in_one_step_equal_per_rate_steps(engine_factory):
    """BASELINE.json configs[3]: streams at different source rates advance in ONE cvad_step (src_rates[]).
    Streams are independent, so the result must be bit-identical to stepping each rate group on its own."""
    rates = np.array([48000, 24000, 16000, 8000, 24000, 48000, 48000, 16000, 8000, 24000] * 7, np.int32)[:67]
    n, T = rates.size, 5
    rng = np.random.default_rng(41)
    n_in = rates * 512 // 16000
    stride = int(T * n_in.max())
    audio = np.zeros((n, stride), np.float32)
    t = np.arange(stride)
    for i in range(n):
        L = T * n_in[i]
        audio[i, :L] = (0.2 * np.sin(2 * np.pi * (180 + 7 * i) * t[:L] / rates[i]) * (1 + 0.5 * np.sin(2 * np.pi * 3 * t[:L] / rates[i]))
                        + 0.02 * rng.standard_normal(L)).astype(np.float32)
    nfr = np.full(n, T, np.int32)
    nfr[3] = 2
    nfr[10] = 0
    eng = engine_factory(128)
    eng.reset()
    eng.configure(enable_denoising=True)
    mixed = eng.step(audio, src_rates=rates, max_frames=T, n_frames=nfr)
    assert mixed.status.sum() == 0
    # reference: one engine step per rate group, same slots
    eng.reset()
    want = np.zeros((n, T), np.float32)
    for r in (8000, 16000, 24000, 48000):
        idx = np.nonzero(rates == r)[0]
        L = T * (r * 512 // 16000)
        res = eng.step(np.ascontiguousarray(audio[idx, :L]), slots=idx, src_rate=r, n_frames=nfr[idx], max_frames=T)
        want[idx] = res.probs
    assert np.array_equal(mixed.probs, want)
    assert mixed.probs[10].sum() == 0 and np.all(mixed.probs[3, 2:] == 0)
    # a second mixed step carries the per-slot state on
    again = eng.step(audio, src_rates=rates, max_frames=T, n_frames=nfr)
    assert again.probs.shape == (n, T)


def test_mixed_source_rates_reject_unknown_rate(engine_factory):
    from real_time_vad.engine.stream_engine import EngineError
    eng = engine_factory(8)
    with pytest.raises(EngineError):
        eng.step(np.zeros((2, 1536), np.float32), src_rates=[48000, 44100], max_frames=1)
