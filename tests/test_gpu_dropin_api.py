"""GPU: the reference-facing API on the CUDA engine vs golden vectors produced by the
reference's unmodified Python (tests/golden/sample_voice.npz, oracle/make_golden.py).

Bars: probabilities within 1e-4 absolute, events at identical frame indices, callback
payloads (WAV bytes) identical."""
import hashlib

import numpy as np
import pytest

from conftest import GOLDEN, synth_streams

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["tc", "fp32", "tc16"])
def math_mode(request, monkeypatch):
    """Every test runs on both arithmetic paths of the v5 engine: tcgen05 tensor cores (BF16 3-way split) and
    packed FP32 FMA.  cvad_create reads CVAD_MATH; engines are created inside the tests."""
    monkeypatch.setenv("CVAD_MATH", request.param)
    from real_time_vad.engine import pool
    for (version, _dev, _path), engines in pool._ENGINES.items():   # engines the process-wide pool already holds
        if version == "v5":
            for pe in engines:
                pe.engine.set_math(request.param)
    return request.param

TOL = 1e-4


def _spy_wrapper(cfg):
    from real_time_vad import VADWrapper
    w = VADWrapper(cfg)
    log = {"events": [], "wavs": [], "cont": 0}
    w.set_callbacks(
        voice_start_callback=lambda: log["events"].append((w._state.total_frames_processed, 1)),
        voice_end_callback=lambda b: (log["events"].append((w._state.total_frames_processed, 2)), log["wavs"].append(b)),
        voice_continue_callback=lambda b: log.__setitem__("cont", log["cont"] + 1))
    return w, log


@pytest.mark.parametrize("mode", ["A", "B", "C"])
def test_vadwrapper_matches_reference_python_on_sample_voice(mode):
    from real_time_vad import SampleRate, VADConfig
    g = np.load(GOLDEN / "sample_voice.npz")
    if mode == "A":
        cfg = VADConfig(sample_rate=SampleRate.SAMPLERATE_16, buffer_size=480, vad_start_probability=0.4,
                        vad_end_probability=0.3, voice_start_frame_count=6, voice_end_frame_count=12)
        q = g["q16k"]
        chunks = [q[i * 480:(i + 1) * 480].astype(np.float32) / 32767.0 for i in range(len(q) // 480)]
    elif mode == "B":
        cfg, chunks = VADConfig(), [g["y16k"]]
    else:
        cfg = VADConfig()
        y = g["y16k"]
        chunks = [y[i * 512:(i + 1) * 512] for i in range(len(y) // 512)]
    w, log = _spy_wrapper(cfg)
    assert w.processor.supports_batched() is True
    probs = []
    for c in chunks:
        before = len(w.processor.voice_probabilities)
        w.process_audio_data(c)
        if mode != "B":
            probs.append(w.processor.voice_probabilities[-1])
    want_p = g[f"{mode}_probs"]
    assert w.get_statistics()["total_frames_processed"] == len(want_p)
    if mode == "B":
        probs = list(w.processor.voice_probabilities)                  # deque(maxlen=100): the last 100
        assert np.abs(np.array(probs) - want_p[-100:]).max() <= TOL
    else:
        assert np.abs(np.array(probs) - want_p).max() <= TOL
    assert log["events"] == [tuple(e) for e in g[f"{mode}_events"].tolist()]
    assert [len(b) for b in log["wavs"]] == g[f"{mode}_wav_len"].tolist()
    assert [hashlib.sha256(b).hexdigest() for b in log["wavs"]] == g[f"{mode}_wav_sha"].tolist()
    assert log["cont"] > 0
    if mode == "A":
        assert [k for _, k in log["events"]] == [1, 2] * 4               # the reference's 4-segment known answer
    w.cleanup()


def test_compat_session_run_equals_fast_path_and_oracle(ref_v5):
    """`ort.InferenceSession.run` (state in / state out, like onnxruntime) vs the batched step."""
    from real_time_vad.core.config import SileroModelVersion
    from real_time_vad.core.silero_model import SileroVADModel
    from conftest import V5_ONNX
    g = np.load(GOLDEN / "v5_frames.npz")
    m = SileroVADModel(str(V5_ONNX), SileroModelVersion.V5)
    assert m.get_model_info()["has_cuda"] is True
    for name in ("zeros", "sine440", "noise"):
        m.reset()
        got = [m.predict(g[f"in_{name}"], 16000) for _ in range(6)]
        assert np.abs(np.array(got) - g[f"p_{name}"]).max() <= TOL
        assert np.abs(m.model_state.state - g[f"state_{name}"]).max() <= 1e-3
    assert m.prediction_count == 18
    with pytest.raises(Exception, match="Model prediction failed"):
        m.predict(g["in_noise"], 8000)                                   # only the 16 kHz branch exists


def test_process_frame_and_batched_paths_interleave_consistently():
    from real_time_vad import VADConfig, VADWrapper
    x = synth_streams(1, 512 * 60, seed=31)[0]
    cfg = VADConfig(vad_start_probability=0.5, vad_end_probability=0.35, voice_start_frame_count=3,
                    voice_end_frame_count=4)
    a, b = VADWrapper(cfg), VADWrapper(cfg)
    ev_a, ev_b = [], []
    a.set_callbacks(voice_start_callback=lambda: ev_a.append(("S", a._state.total_frames_processed)),
                    voice_end_callback=lambda w_: ev_a.append(("E", a._state.total_frames_processed, len(w_))))
    b.set_callbacks(voice_start_callback=lambda: ev_b.append(("S", b._state.total_frames_processed)),
                    voice_end_callback=lambda w_: ev_b.append(("E", b._state.total_frames_processed, len(w_))))
    a.process_audio_data(x)                                              # one batched call, hop 256
    # b: alternate between the per-frame compat path and small batched calls over the same frames
    pos, frames_done, k = 0, 0, 0
    n_total = (len(x) - 512) // 256 + 1
    while frames_done < n_total:
        if k % 2 == 0:
            r = b.processor.process_frame(x[pos:pos + 512].copy())
            b._handle_callbacks(r)
            b._state.total_frames_processed += 1
            pos += 256
            frames_done += 1
        else:
            take = min(5, n_total - frames_done)
            b.process_audio_data(x[pos:pos + 512 + (take - 1) * 256])
            pos += take * 256
            frames_done += take
        k += 1
    assert ev_a == ev_b and len(ev_a) >= 2
    pa, pb = np.array(a.processor.voice_probabilities), np.array(b.processor.voice_probabilities)
    assert np.abs(pa - pb).max() <= 2e-6


def test_batched_manager_equals_per_stream_wrappers(ref_v5, ref_lib):
    from real_time_vad import BatchedVADManager, VADConfig
    from vad_oracle import sm_run_c
    n, T = 37, 90
    audio = synth_streams(n, 512 * T, seed=41)
    cfgs = [VADConfig(vad_start_probability=(0.4, 0.6)[s % 2], vad_end_probability=0.3,
                      voice_start_frame_count=(3, 5)[s % 2], voice_end_frame_count=(4, 9)[(s // 2) % 2],
                      enable_denoising=bool(s % 3)) for s in range(n)]
    mgr = BatchedVADManager(max_streams=64, frame_len=512, hop=512)
    logs = {}
    ids = []
    for s in range(n):
        logs[s] = []
        ids.append(mgr.open_stream(
            cfgs[s], on_voice_start=(lambda s=s: logs[s].append(("S",))),
            on_voice_end=(lambda b, s=s: logs[s].append(("E", len(b)))) if s % 2 == 0 else None))
    rng = np.random.default_rng(5)
    all_events, probs = [], {s: [] for s in range(n)}
    pos = np.zeros(n, int)
    while (pos < 512 * T).any():
        for s in range(n):                                               # ragged, jittered arrival
            if pos[s] < 512 * T:
                k = int(rng.integers(100, 1500))
                mgr.push(ids[s], audio[s, pos[s]:pos[s] + k])
                pos[s] += k
        out = mgr.step()
        all_events += out.events
        for sid, p in out.probabilities.items():
            probs[ids.index(sid)].append(p)
    n_ev = 0
    for s in range(n):
        got = np.concatenate(probs[s])
        assert len(got) == T
        want, _, _ = ref_v5.run(audio[s:s + 1], T, denoise=cfgs[s].enable_denoising)
        assert np.abs(got - want[0]).max() <= TOL
        c = cfgs[s]
        fl, _ = sm_run_c(ref_lib, want[0], c.vad_start_probability, c.vad_end_probability, 0.8, 0.95,
                         c.voice_start_frame_count, c.voice_end_frame_count)
        want_ev = [(int(j), "start") for j in np.flatnonzero(fl & 1)] + [(int(j), "end") for j in np.flatnonzero(fl & 2)]
        got_ev = [(e.frame_index, e.kind) for e in all_events if e.stream_id == ids[s]]
        assert sorted(got_ev) == sorted(want_ev), s
        n_ev += len(got_ev)
        assert [e[0] for e in logs[s]].count("S") == sum(k == "start" for _, k in got_ev)
        if s % 2 == 0:
            assert [e[0] for e in logs[s]].count("E") == sum(k == "end" for _, k in got_ev)
    assert n_ev > 20
    # WAV payload of the first ended segment of an even stream == what a VADWrapper would deliver
    from real_time_vad import VADWrapper
    s = next(s for s in range(0, n, 2) if any(e[0] == "E" for e in logs[s]))
    c = cfgs[s].model_copy(update={"buffer_size": 512})
    w = VADWrapper(c)
    w._DEFAULT_FRAME_OVERLAP_RATIO = 1.0                                   # hop == frame, like the manager
    wavs = []
    w.set_callbacks(voice_end_callback=lambda b: wavs.append(len(b)))
    w.process_audio_data(audio[s])
    assert wavs == [e[1] for e in logs[s] if e[0] == "E"]
    mgr.close_stream(ids[0])
    assert ids[0] not in mgr.open_streams
    mgr.close()


def test_manager_rejects_nonfinite_push_and_bad_config():
    from real_time_vad import BatchedVADManager, SampleRate, VADConfig
    from real_time_vad.core.exceptions import AudioProcessingError, ConfigurationError
    mgr = BatchedVADManager(max_streams=8)
    sid = mgr.open_stream()
    with pytest.raises(AudioProcessingError, match="infinite or NaN"):
        mgr.push(sid, np.array([0.1, np.inf], np.float32))
    with pytest.raises(AudioProcessingError, match="empty"):
        mgr.push(sid, np.zeros(0, np.float32))
    with pytest.raises(ConfigurationError):
        mgr.open_stream(VADConfig(sample_rate=SampleRate.SAMPLERATE_8))
    mgr.push(sid, np.zeros(100, np.float32))
    assert mgr.step().frames == 0                                         # less than one frame buffered
    mgr.close()


def test_manager_push_many_lockstep_matches_single_pushes(ref_v5):
    """Vectorised producer path (push_many) == per-stream push, incl. compaction of leftovers."""
    from real_time_vad import BatchedVADManager, VADConfig
    n, T = 70, 25
    audio = synth_streams(n, 512 * T + 100, seed=51)
    cfg = VADConfig(enable_denoising=True, vad_start_probability=0.5, vad_end_probability=0.35,
                    voice_start_frame_count=2, voice_end_frame_count=3)
    a = BatchedVADManager(max_streams=128, frame_len=512, hop=256)
    b = BatchedVADManager(max_streams=128, frame_len=512, hop=256)
    ia = [a.open_stream(cfg) for _ in range(n)]
    ib = [b.open_stream(cfg) for _ in range(n)]
    pa, pb, ea, eb = [], [], [], []
    for lo in range(0, audio.shape[1], 700):
        chunk = audio[:, lo:lo + 700]
        a.push_many(ia, chunk)
        for k, sid in enumerate(ib):
            b.push(sid, chunk[k])
        oa, ob = a.step(), b.step()
        pa.append(oa.probs[:, :int(oa.counts.max())] if oa.counts.size else np.zeros((n, 0), np.float32))
        pb.append(ob.probs[:, :int(ob.counts.max())] if ob.counts.size else np.zeros((n, 0), np.float32))
        ea += [(e.stream_id, e.kind, e.frame_index) for e in oa.events]
        eb += [(e.stream_id, e.kind, e.frame_index) for e in ob.events]
    pa, pb = np.concatenate(pa, axis=1), np.concatenate(pb, axis=1)
    assert np.array_equal(pa, pb) and ea == eb and len(ea) > 10
    Tn = (audio.shape[1] - 512) // 256 + 1
    assert pa.shape == (n, Tn)
    want, _, _ = ref_v5.run(audio, Tn, hop=256, denoise=True)
    assert np.abs(pa - want).max() <= TOL
    assert a.is_voice_active(ia[0]) == bool((a.engine.get_state(ia[0])[2][0]))
    a.close()
    b.close()


def test_manager_mixed_source_rates_one_step_equals_per_rate_managers():
    """BatchedVADManager(source_rate=None): every stream has its own VADConfig.sample_rate and all of them advance
    in one engine step; probabilities, events and voice-end payloads equal those of one manager per rate."""
    from real_time_vad import BatchedVADManager, SampleRate, VADConfig
    from scipy import signal
    rates = {8000: SampleRate.SAMPLERATE_8, 16000: SampleRate.SAMPLERATE_16, 24000: SampleRate.SAMPLERATE_24,
             48000: SampleRate.SAMPLERATE_48}
    T = 40
    base = synth_streams(8, 512 * T, seed=77)
    plan = [48000, 16000, 24000, 8000, 48000, 24000, 16000, 8000]

    def cfg(rate):
        return VADConfig(sample_rate=rates[rate], enable_denoising=True, vad_start_probability=0.5,
                         vad_end_probability=0.35, voice_start_frame_count=2, voice_end_frame_count=3)

    audio = [signal.resample(base[k], plan[k] * 512 // 16000 * T).astype(np.float32) for k in range(8)]
    mixed = BatchedVADManager(max_streams=16, source_rate=None)
    ends_m = {k: [] for k in range(8)}
    ids_m = [mixed.open_stream(cfg(plan[k]), on_voice_end=ends_m[k].append) for k in range(8)]
    solo = {r: BatchedVADManager(max_streams=16, source_rate=r) for r in set(plan)}
    ends_s = {k: [] for k in range(8)}
    ids_s = [solo[plan[k]].open_stream(cfg(plan[k]), on_voice_end=ends_s[k].append) for k in range(8)]
    probs_m = {k: [] for k in range(8)}
    probs_s = {k: [] for k in range(8)}
    ev_m, ev_s = [], []
    for step in range(5):                                   # ragged arrival: 8 chunks per push
        for k in range(8):
            n_in = plan[k] * 512 // 16000
            piece = audio[k][step * 8 * n_in:(step + 1) * 8 * n_in]
            mixed.push(ids_m[k], piece)
            solo[plan[k]].push(ids_s[k], piece)
        out = mixed.step()
        for sid, p in out.probabilities.items():
            probs_m[ids_m.index(sid)].append(p)
        ev_m += [(ids_m.index(e.stream_id), e.kind, e.frame_index) for e in out.events]
        for r, mgr in solo.items():
            o = mgr.step()
            for sid, p in o.probabilities.items():
                k = [kk for kk in range(8) if plan[kk] == r and ids_s[kk] == sid][0]
                probs_s[k].append(p)
            ev_s += [([kk for kk in range(8) if plan[kk] == r and ids_s[kk] == e.stream_id][0], e.kind, e.frame_index)
                     for e in o.events]
    for k in range(8):
        assert np.array_equal(np.concatenate(probs_m[k]), np.concatenate(probs_s[k])), k
        assert len(np.concatenate(probs_m[k])) == T
        assert ends_m[k] == ends_s[k]
    assert sorted(ev_m) == sorted(ev_s) and len(ev_m) > 0
    mixed.close()
    for mgr in solo.values():
        mgr.close()


# ---------------------------------------------------------------------------------------------------------------
# The rest of the reference's examples/audios and its v4 model selection, against the reference's own Python
# ---------------------------------------------------------------------------------------------------------------

def _run_golden(g, mode, cfg, chunks, full_probs_from=None):
    w, log = _spy_wrapper(cfg)
    assert w.processor.supports_batched() is True
    probs = []
    for c in chunks:
        w.process_audio_data(c)
        probs.append(w.processor.voice_probabilities[-1])
    want_p = g[f"{mode}_probs"]
    assert w.get_statistics()["total_frames_processed"] == len(want_p)
    if len(chunks) == 1:                                                 # one call: the deque keeps the last 100
        got = np.array(w.processor.voice_probabilities)
        assert np.abs(got - want_p[-len(got):]).max() <= TOL
    else:
        assert np.abs(np.array(probs) - want_p).max() <= TOL
    assert log["events"] == [tuple(e) for e in g[f"{mode}_events"].tolist()]
    assert [len(b) for b in log["wavs"]] == g[f"{mode}_wav_len"].tolist()
    assert [hashlib.sha256(b).hexdigest() for b in log["wavs"]] == g[f"{mode}_wav_sha"].tolist()
    w.cleanup()
    return log


@pytest.mark.parametrize("mode", ["A", "B"])
def test_vadwrapper_v4_matches_reference_python_on_sample_voice(mode, math_mode):
    """VADConfig(model_version=V4) (config.py:242, silero_model.py:369-376): events, probabilities and the SHA-256 of
    every WAV payload equal what the reference's unmodified Python produced on SampleVoiceMono.wav
    (tests/golden/sample_voice_v4.npz).  The v4 engine runs its default build (CVAD_MATH_FFT) whatever the v5 math is."""
    if math_mode != "tc16":
        pytest.skip("v4 engines do not follow the v5 math parameter; run once")
    from real_time_vad import SampleRate, SileroModelVersion, VADConfig
    g4 = np.load(GOLDEN / "sample_voice_v4.npz")
    g = np.load(GOLDEN / "sample_voice.npz")
    if mode == "A":
        cfg = VADConfig(model_version=SileroModelVersion.V4, sample_rate=SampleRate.SAMPLERATE_16, buffer_size=480,
                        vad_start_probability=0.4, vad_end_probability=0.3, voice_start_frame_count=6, voice_end_frame_count=12)
        q = g["q16k"]
        chunks = [q[i * 480:(i + 1) * 480].astype(np.float32) / 32767.0 for i in range(len(q) // 480)]
    else:
        cfg, chunks = VADConfig(model_version=SileroModelVersion.V4), [g["y16k"]]
    log = _run_golden(g4, mode, cfg, chunks)
    assert [k for _, k in log["events"]] == [1, 2] * 4                   # v4 also finds the 4 segments


def test_stereo_sample_file_through_vadwrapper_and_on_the_device(ref_v5):
    """examples/audios/SampleVoiceStereo.wav: both channels equal SampleVoiceMono.wav sample for sample (asserted when
    the golden was made), so the stereo input is rebuilt from the committed mono file.  (1) VADWrapper fed the 2-D
    [samples, 2] array reproduces the reference's events / probabilities / WAV hashes; (2) the same array through the
    engine's own down-mix (cvad_step_args.channels = 2) gives the same probabilities as the mono stream."""
    from real_time_vad import VADConfig
    from real_time_vad.engine.stream_engine import StreamEngine
    gs = np.load(GOLDEN / "sample_voice_stereo.npz")
    g = np.load(GOLDEN / "sample_voice.npz")
    y = g["y16k"]
    _run_golden(gs, "file", VADConfig(), [np.stack([y, y], axis=1)])
    yd = np.stack([y, gs["diff_right"]], axis=1).astype(np.float32)      # channels that differ: the mean matters
    _run_golden(gs, "diff", VADConfig(vad_start_probability=0.4, vad_end_probability=0.3, voice_start_frame_count=6,
                                      voice_end_frame_count=12), [yd])
    # device down-mix: [1, samples, 2] through the C ABI, hop 256 like the wrapper
    eng = StreamEngine("v5", max_streams=4)
    eng.configure(enable_denoising=True)
    T = (len(y) - 512) // 256 + 1
    r = eng.step(yd[None], hop=256)
    assert r.probs.shape == (1, T)
    assert np.abs(r.probs[0] - gs["diff_probs"]).max() <= TOL
    eng.reset()
    r2 = eng.step(np.mean(yd, axis=1)[None].astype(np.float32), hop=256)   # the host's np.mean, then mono
    assert np.array_equal(r.probs, r2.probs)                              # the device's mean IS numpy's float32 mean
    eng.close()


def test_device_downmix_pcm16_three_channels_ragged(engine_factory, ref_v5):
    """Row M on the device beyond the sample file: 3 interleaved int16 channels, several streams, several frames."""
    from real_time_vad.engine import capi
    eng = engine_factory(64)
    eng.reset()
    eng.configure(enable_denoising=True)
    n, T, C = 21, 7, 3
    rng = np.random.default_rng(5)
    mono = synth_streams(n, 512 * T, seed=12)
    x = np.stack([mono * g + 0.01 * rng.standard_normal(mono.shape) for g in (1.0, 0.5, -0.25)], axis=2)
    q = np.clip(np.round(x * 32767.0), -32768, 32767).astype(np.int16)
    xf = q.astype(np.float32) / np.float32(32767.0)
    want_in = np.mean(xf, axis=2)                                          # AudioUtils.convert_to_mono, float32
    assert want_in.dtype == np.float32
    want, _, _ = ref_v5.run(want_in, T, denoise=True)
    r = eng.step(q, pcm_format=capi.PCM_S16_32767)
    assert np.abs(r.probs - want).max() <= TOL


def test_callback_that_raises_mid_call_leaves_the_stream_where_the_reference_does():
    """vad_wrapper.py:470-476: a callback exception aborts the call at that frame; frames after it never ran.  The
    batched path has already stepped them on the device, so it must put the stream back: LSTM state, counters and the
    probability history after the call equal those of a wrapper that was fed exactly the frames up to the failing one."""
    from real_time_vad import VADConfig, VADWrapper
    from real_time_vad.core.exceptions import AudioProcessingError
    x = synth_streams(1, 512 * 40, seed=31)[0]
    cfg = VADConfig(vad_start_probability=0.5, vad_end_probability=0.35, voice_start_frame_count=3, voice_end_frame_count=4)
    a, b = VADWrapper(cfg), VADWrapper(cfg)
    seen = {"n": 0}

    def boom(_pcm):
        seen["n"] += 1
        if seen["n"] == 5:
            raise RuntimeError("consumer failed")
    a.set_callbacks(voice_continue_callback=boom)
    with pytest.raises(AudioProcessingError):
        a.process_audio_data(x)
    done = a._state.total_frames_processed + 1            # the failing frame ran in the reference too (callbacks come after it)
    assert 5 <= done < (len(x) - 512) // 256 + 1
    b.process_audio_data(x[:512 + (done - 1) * 256])       # exactly those frames, no callback trouble
    pa, pb = a.processor, b.processor
    assert np.array_equal(pa.model.model_state.state, pb.model.model_state.state)
    assert (pa.is_voice_active, pa.voice_start_frame_count, pa.voice_end_frame_count) == \
           (pb.is_voice_active, pb.voice_start_frame_count, pb.voice_end_frame_count)
    assert list(pa.voice_probabilities) == list(pb.voice_probabilities)
    assert pa.model.prediction_count == pb.model.prediction_count == done
    # and the stream goes on from there like the other one
    rest = x[done * 256:]
    a.set_callbacks(voice_continue_callback=lambda _b: None)
    a.process_audio_data(rest)
    b.process_audio_data(rest)
    assert np.array_equal(a.processor.model.model_state.state, b.processor.model.model_state.state)


@pytest.mark.parametrize("rate", [8000, 48000])
def test_vadwrapper_input_sample_rate_opt_in_runs_the_gpu_resampler(ref_v5, ref_lib, rate):
    """VADWrapper.process_audio_data(audio, input_sample_rate=rate): the wrapper's route to the GPU resampler (the reference
    leaves `auto_convert_sample_rate` a placeholder, vad_wrapper.py:619-624).  Chunks of 512*rate/16000 source samples,
    each resampled to one model frame on the device = AudioUtils.resample_audio per chunk, then the model: probabilities
    within 1e-4 of the oracle on the scipy-resampled audio, events at the oracle's frames, WAV payloads = the host-
    resampled, gated segment; the same stream through BatchedVADManager(source_rate=rate) agrees to 5e-5; calls carry the state across; a rate the engine does not know is an AudioProcessingError."""
    from scipy import signal

    from real_time_vad import BatchedVADManager, VADConfig
    from real_time_vad.core.exceptions import AudioProcessingError
    from vad_oracle import resample_chunks, sm_run_c
    T, n_in = 60, 512 * rate // 16000
    cand = synth_streams(12, 512 * T, seed=77)
    cand_up = signal.resample(cand, n_in * T, axis=1).astype(np.float32)
    cand_p, _, _ = ref_v5.run(resample_chunks(cand_up, rate), T, denoise=True)
    n_events = [int(((sm_run_c(ref_lib, cand_p[s], 0.5, 0.35, 0.8, 0.95, 2, 3)[0] & 3) != 0).sum()) for s in range(12)]
    x = cand_up[int(np.argmax(n_events))]                              # the candidate stream with the most start / end events
    cfg = VADConfig(vad_start_probability=0.5, vad_end_probability=0.35, voice_start_frame_count=2,
                    voice_end_frame_count=3, enable_denoising=True, buffer_size=512)
    w, log = _spy_wrapper(cfg)
    w.process_audio_data(x[:n_in * 25], input_sample_rate=rate)
    w.process_audio_data(x[n_in * 25:], input_sample_rate=rate)
    got = np.array(w.processor.voice_probabilities, np.float32)[-T:]
    y16 = resample_chunks(x[None, :], rate)
    want, _, _ = ref_v5.run(y16, T, denoise=True)
    assert got.shape == (T,) and np.abs(got - want[0]).max() <= TOL
    fl, _ = sm_run_c(ref_lib, want[0], 0.5, 0.35, 0.8, 0.95, 2, 3)
    want_ev = sorted([(int(j) + 1, 1) for j in np.flatnonzero(fl & 1)] + [(int(j) + 1, 2) for j in np.flatnonzero(fl & 2)])
    near = (np.abs(want[0] - 0.5) <= TOL).any() or (np.abs(want[0] - 0.35) <= TOL).any()
    if not near:
        assert sorted((f + 1, k) for f, k in log["events"]) == want_ev and len(want_ev) >= 2
    # the same stream through the manager's source-rate path: same kernels, same bits
    mgr = BatchedVADManager(max_streams=4, source_rate=rate)
    sid = mgr.open_stream(VADConfig(vad_start_probability=0.5, vad_end_probability=0.35, voice_start_frame_count=2,
                                    voice_end_frame_count=3, enable_denoising=True, sample_rate=rate))
    mgr.push(sid, x)
    mp = []
    while len(mp) < T:
        out = mgr.step()
        assert sid in out.probabilities
        mp += list(out.probabilities[sid])
    # (the manager steps at most 8 frames at a time and runs its last frames through the fused one-frame kernel: same
    # arithmetic builds, different kernel forms -- agreement well inside the bar, not bit for bit)
    assert np.abs(np.asarray(mp, np.float32) - got).max() <= 5e-5
    mgr.close()
    with pytest.raises(AudioProcessingError):
        w.process_audio_data(x[:n_in * 4], input_sample_rate=44100)


def test_wrapper_slot_cache_follows_reset_and_config_changes():
    """The wrapper's fast path keeps thresholds and LSTM state on its engine slot between calls and uploads them only
    when the host's copies differ (core/silero_model.py: process_audio_batched.run).  A reset, a state assigned by hand
    and a new configuration must each reach the device: every run below equals a fresh wrapper given the same history."""
    from real_time_vad import VADConfig
    audio = synth_streams(1, 512 * 40, seed=91)[0]
    cfg_a = VADConfig(vad_start_probability=0.5, vad_end_probability=0.35, voice_start_frame_count=2, voice_end_frame_count=3)
    cfg_b = VADConfig(vad_start_probability=0.8, vad_end_probability=0.6, voice_start_frame_count=4, voice_end_frame_count=2,
                      enable_denoising=False)

    def run(w, log, lo, hi):
        n0 = len(log["events"])
        for i in range(lo, hi, 2048):                                  # several calls: the cached path
            w.process_audio_data(audio[i:min(i + 2048, hi)])
        return list(w.processor.voice_probabilities), log["events"][n0:]

    w, log = _spy_wrapper(cfg_a)
    first = run(w, log, 0, 512 * 20)
    w.reset()
    again = run(w, log, 0, 512 * 20)
    assert np.array_equal(first[0], again[0])                          # reset reached the slot: same bits from scratch
    assert [k for _, k in first[1]] == [k for _, k in again[1]]
    cont = run(w, log, 512 * 20, 512 * 40)                             # carries on from the cached state ...
    ref, rlog = _spy_wrapper(cfg_a)
    run(ref, rlog, 0, 512 * 20)
    want = run(ref, rlog, 512 * 20, 512 * 40)
    assert np.array_equal(cont[0], want[0])                            # ... like a wrapper that was never reset
    w.update_config(cfg_b)                                             # new thresholds, gate off: re-configured slot
    got_b = run(w, log, 0, 512 * 20)
    fresh, flog = _spy_wrapper(cfg_b)
    want_b = run(fresh, flog, 0, 512 * 20)
    assert np.array_equal(got_b[0], want_b[0]) and [k for _, k in got_b[1]] == [k for _, k in want_b[1]]
    # a state assigned by hand (what a caller of the reference may do) is uploaded before the next call
    ms = w.processor.model.model_state
    ms.state = np.zeros_like(ms.state)
    w.processor.is_voice_active = False
    w.processor.voice_start_frame_count = 0
    w.processor.voice_end_frame_count = 0
    w.processor.voice_probabilities.clear()
    manual = run(w, log, 0, 512 * 20)
    assert np.array_equal(manual[0], want_b[0])
