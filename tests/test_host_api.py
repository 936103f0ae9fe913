"""CPU: host-side mirror of the reference API (names, argument meaning, error behaviour).

The model is replaced by a scripted session at the same seam the reference's tests use
(`real_time_vad.core.silero_model.ort.InferenceSession`), which drives the per-frame
COMPATIBILITY path; the fast path needs the GPU and is covered by tests/test_gpu_*.py.
"""
import io
import wave
from unittest.mock import patch

import numpy as np
import pytest

from real_time_vad import (AsyncVADWrapper, AudioUtils, SampleRate, SileroModelVersion, VADConfig, VADError,
                           VADWrapper, WAVWriter)
from real_time_vad.core.exceptions import (AudioProcessingError, CallbackError, ConfigurationError,
                                           ModelInitializationError, ModelNotFoundError)
from real_time_vad.core.silero_model import ModelState, ProcessingResult, SileroVADModel, VADProcessor


class ScriptedSession:
    """onnxruntime-shaped test double: returns the next scripted probability per run()."""

    def __init__(self, probs, version="v5"):
        self.probs = list(probs)
        self.calls = []
        self.version = version

    def get_inputs(self):
        return [object()] * (3 if self.version == "v5" else 4)

    def get_outputs(self):
        return [object()] * (2 if self.version == "v5" else 3)

    def get_providers(self):
        return ["CPUExecutionProvider"]

    def run(self, names, feeds):
        self.calls.append({k: np.array(v, copy=True) for k, v in feeds.items()})
        p = self.probs.pop(0) if self.probs else 0.0
        out = np.array([[p]], np.float32)
        if self.version == "v5":
            return [out, feeds["state"] + 1.0]
        return [out, feeds["h"] + 1.0, feeds["c"] + 2.0]


def make_wrapper(probs, **cfg):
    sess = ScriptedSession(probs)
    with patch("real_time_vad.core.silero_model.ort.InferenceSession", return_value=sess):
        w = VADWrapper(VADConfig(**cfg))
    return w, sess


def test_public_surface_matches_reference_names():
    import real_time_vad
    for name in ("VADWrapper", "AsyncVADWrapper", "VADConfig", "SampleRate", "SileroModelVersion", "VADError",
                 "ModelNotFoundError", "ConfigurationError", "AudioUtils", "WAVWriter", "BatchedVADManager"):
        assert hasattr(real_time_vad, name)
    for meth in ("process_audio_data", "process_audio_data_with_buffer", "set_sample_rate", "set_silero_model",
                 "set_thresholds", "set_callbacks", "reset", "cleanup", "get_statistics", "get_config",
                 "update_config", "is_voice_active", "get_last_error", "get_last_error_details"):
        assert callable(getattr(VADWrapper, meth))
    for meth in ("process_audio_data_async", "set_async_callbacks", "set_thresholds_async", "reset_async",
                 "get_statistics_async", "is_voice_active_async", "update_config_async", "acleanup"):
        assert callable(getattr(AsyncVADWrapper, meth))


def test_config_defaults_and_loaders(tmp_path, monkeypatch):
    c = VADConfig()
    assert (c.sample_rate, c.model_version, c.buffer_size) == (SampleRate.SAMPLERATE_16, SileroModelVersion.V5, 512)
    assert (c.vad_start_probability, c.vad_end_probability, c.voice_start_frame_count,
            c.voice_end_frame_count, c.enable_denoising) == (0.7, 0.7, 10, 50, True)
    assert c.get_model_filename() == "silero_vad_v5.onnx"
    assert VADConfig(model_version=SileroModelVersion.V4).get_model_filename() == "silero_vad.onnx"
    assert c.get_frame_duration_ms() == 32.0
    with pytest.raises(ValueError):
        VADConfig(buffer_size=100)
    with pytest.raises(ValueError, match="Model path does not exist"):
        VADConfig(model_path=tmp_path / "nope")
    d = VADConfig.from_dict({"sample_rate": 8000, "model_version": "V4", "voice_end_frame_count": 7})
    assert d.sample_rate == SampleRate.SAMPLERATE_8 and d.model_version == SileroModelVersion.V4
    y = tmp_path / "c.yaml"
    d.to_yaml(y)
    assert VADConfig.from_yaml(y) == d
    with pytest.raises(FileNotFoundError, match="Configuration file not found"):
        VADConfig.from_yaml(tmp_path / "missing.yaml")
    monkeypatch.setenv("VAD_START_PROBABILITY", "0.55")
    monkeypatch.setenv("VAD_ENABLE_DENOISING", "off")
    e = VADConfig.from_env()
    assert e.vad_start_probability == 0.55 and e.enable_denoising is False


def test_exception_family_codes_and_messages():
    assert str(ModelNotFoundError("/x")) == "[MODEL_NOT_FOUND] Silero model not found at path: /x"
    assert str(ConfigurationError("p", "v")) == "[CONFIGURATION_ERROR] Invalid configuration for parameter 'p': v"
    assert str(AudioProcessingError("boom")) == "[AUDIO_PROCESSING_ERROR] boom"
    assert ModelInitializationError("v5").model_version == "v5"
    ce = CallbackError("voice_start", RuntimeError("x"))
    assert ce.callback_name == "voice_start" and "Error in callback 'voice_start': x" in str(ce)
    assert all(issubclass(c, VADError) for c in (ModelNotFoundError, ConfigurationError, AudioProcessingError,
                                                 ModelInitializationError, CallbackError))


def test_compat_path_feeds_model_like_the_reference_and_fires_callbacks_in_order():
    # N_s=3, N_e=2; frame = 512, hop = 256 -> 9 frames from 2560 samples
    probs = [0.1, 0.8, 0.8, 0.8, 0.9, 0.2, 0.2, 0.1, 0.1]
    w, sess = make_wrapper(probs, vad_start_probability=0.7, vad_end_probability=0.3, voice_start_frame_count=3,
                           voice_end_frame_count=2)
    log = []
    w.set_callbacks(voice_start_callback=lambda: log.append("start"),
                    voice_end_callback=lambda b: log.append(("end", len(b))),
                    voice_continue_callback=lambda b: log.append(("cont", len(b))))
    x = (0.5 * np.sin(np.arange(2560) / 5.0)).astype(np.float32)
    x[:10] = 0.005                                        # below the noise gate -> must reach the model as 0
    w.process_audio_data(x)
    assert len(sess.calls) == 9
    first = sess.calls[0]
    assert set(first) == {"input", "state", "sr"} and first["input"].shape == (1, 512)
    assert first["sr"].dtype == np.int64 and first["sr"].tolist() == [16000]
    assert not first["input"][0, :10].any()
    assert np.array_equal(first["input"][0], np.where(np.abs(x[:512]) > 0.01, x[:512], 0))
    assert np.array_equal(sess.calls[1]["input"][0], np.where(np.abs(x[256:768]) > 0.01, x[256:768], 0))
    assert sess.calls[3]["state"].flat[0] == 3.0          # state output fed back each frame
    # start on the 3rd hit (frame 3); frame 4 continues; frames 5,6 are quiet -> end on frame 6
    kinds = [e if isinstance(e, str) else e[0] for e in log]
    assert kinds == ["start", "cont", "cont", "end", "cont"]
    # WAV = 44-byte header + int16 of the 3 pre-roll frames + frames 4,5,6
    assert log[3] == ("end", 44 + 2 * 512 * 6)
    assert log[1] == ("cont", 4 * 512)
    st = w.get_statistics()
    assert st["total_frames_processed"] == 9 and st["is_voice_active"] is False and st["has_callbacks"]
    assert w.processor.model.prediction_count == 9


def test_error_messages_match_reference_prefixes():
    w, _ = make_wrapper([0.5])
    with pytest.raises(AudioProcessingError, match="Audio data cannot be empty"):
        w.process_audio_data([])
    with pytest.raises(AudioProcessingError, match="Unsupported audio data type"):
        w.process_audio_data("abc")
    with pytest.raises(AudioProcessingError, match="infinite or NaN"):
        w.process_audio_data(np.array([0.0, np.nan] * 300, np.float32))
    assert "infinite or NaN" in w.get_last_error()
    with pytest.raises(AudioProcessingError, match="Frame processing failed"):
        w.process_audio_data(np.zeros(100, np.float32))   # 1..255 samples: negative frame count
    w.process_audio_data(np.zeros(300, np.float32))        # 256..511 samples: zero frames, no error
    with pytest.raises(AudioProcessingError, match="Count must be non-negative"):
        w.process_audio_data_with_buffer(np.zeros(10, np.float32), -1)
    with pytest.raises(AudioProcessingError, match="Count .* exceeds buffer size"):
        w.process_audio_data_with_buffer(np.zeros(10, np.float32), 11)
    with pytest.raises(ConfigurationError):
        w.set_thresholds(vad_start_probability=0.05)
    with pytest.raises(ConfigurationError):
        w.set_sample_rate(16000)
    w.cleanup()
    with pytest.raises(VADError, match="VAD processor not initialized"):
        w.process_audio_data(np.zeros(512, np.float32))


def test_callback_exception_becomes_callback_error_and_aborts_the_call():
    w, sess = make_wrapper([0.9] * 10, voice_start_frame_count=1)

    def boom():
        raise RuntimeError("cb failed")
    w.set_callbacks(voice_start_callback=boom)
    with pytest.raises(AudioProcessingError, match="Error in callback 'voice_start'"):
        w.process_audio_data(np.full(2048, 0.5, np.float32))
    assert len(sess.calls) == 1


def test_prediction_failure_and_bad_probability():
    sess = ScriptedSession([1.5])
    with patch("real_time_vad.core.silero_model.ort.InferenceSession", return_value=sess):
        m = SileroVADModel(__file__.replace(".py", ".onnx") if False else _junk_model(), SileroModelVersion.V5)
    with pytest.raises(AudioProcessingError, match="Probability extraction failed"):
        m.predict(np.zeros(512, np.float32), 16000)
    sess.run = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("ORT down"))
    with pytest.raises(AudioProcessingError, match="Model prediction failed"):
        m.predict(np.zeros(512, np.float32), 16000)
    bad = ScriptedSession([0.5], version="v4")
    with patch("real_time_vad.core.silero_model.ort.InferenceSession", return_value=bad):
        with pytest.raises(ModelInitializationError, match="Model signature validation failed"):
            SileroVADModel(_junk_model(), SileroModelVersion.V5)
    with pytest.raises(Exception):
        ModelState(state=np.zeros(3, np.float64))


_JUNK = {}


def _junk_model():
    import tempfile
    if "p" not in _JUNK:
        d = tempfile.mkdtemp()
        _JUNK["p"] = d + "/silero_vad_v5.onnx"
        open(_JUNK["p"], "wb").write(b"not a model")
    return _JUNK["p"]


def test_prepare_audio_input_pads_and_truncates():
    sess = ScriptedSession([0.1, 0.1])
    with patch("real_time_vad.core.silero_model.ort.InferenceSession", return_value=sess):
        m = SileroVADModel(_junk_model(), SileroModelVersion.V5)
    m.predict(np.ones(480, np.float32), 16000)
    m.predict(np.ones(700, np.float32), 16000)
    a, b = sess.calls[0]["input"], sess.calls[1]["input"]
    assert a.shape == b.shape == (1, 512) and a[0, :480].all() and not a[0, 480:].any() and b.all()


def test_wav_writer_bytes_are_a_valid_pcm_file():
    x = np.array([0.0, 0.5, -0.5, 1.0, -1.0, 1.5], np.float32)
    b = WAVWriter(16000, 16, 1).write_wav_data(x)
    with wave.open(io.BytesIO(b)) as wf:
        assert (wf.getnchannels(), wf.getsampwidth(), wf.getframerate(), wf.getnframes()) == (1, 2, 16000, 6)
        pcm = np.frombuffer(wf.readframes(6), np.int16)
    assert pcm.tolist() == [0, 16383, -16383, 32767, -32767, 32767]
    assert len(WAVWriter(16000, 32, 1).write_wav_data(x)) == 44 + 24
    with pytest.raises(ValueError):
        WAVWriter(16000, 24, 1)


def test_audio_utils_hot_path_members():
    x = np.array([0.005, -0.02, 0.01, 0.0100001, -0.3], np.float32)
    assert AudioUtils.denoise_audio(x).tolist() == [0.0, x[1], 0.0, x[3], x[4]]
    assert AudioUtils.convert_to_mono(np.array([[1.0, 3.0], [2.0, 4.0]])).tolist() == [2.0, 3.0]
    with pytest.raises(AudioProcessingError, match="Audio data is empty"):
        AudioUtils.validate_audio_data(np.zeros(0))
    y = AudioUtils.resample_audio(np.sin(np.arange(1536) / 20).astype(np.float32), 48000, 16000)
    assert y.shape == (512,) and y.dtype == np.float32
    assert AudioUtils.resample_audio(x, 16000, 16000) is x
    assert AudioUtils.pcm_to_float32(np.array([16384], np.int16).tobytes()).tolist() == [0.5]


def test_async_wrapper_dispatches_coroutine_callbacks():
    import asyncio
    sess = ScriptedSession([0.9, 0.9, 0.1, 0.1, 0.1])
    with patch("real_time_vad.core.silero_model.ort.InferenceSession", return_value=sess):
        aw = AsyncVADWrapper(VADConfig(voice_start_frame_count=2, voice_end_frame_count=2,
                                       vad_end_probability=0.3))
    seen = []

    async def main():
        async def on_start():
            seen.append("start")

        async def on_end(b):
            seen.append(("end", len(b)))
        aw.set_async_callbacks(voice_start_callback=on_start, voice_end_callback=on_end)
        await aw.process_audio_data_async(np.full(512 + 4 * 256, 0.4, np.float32))
        await asyncio.sleep(0.05)
        assert await aw.is_voice_active_async() is False
    asyncio.run(main())
    assert seen[0] == "start" and seen[1][0] == "end"
    aw.cleanup()


def test_v4_weight_blobs_of_both_branches():
    """canonical_blob_v4 lays out the 16 kHz branch or the 8 kHz sub-model (`model_8k.*`) in the same order."""
    from conftest import V4_ONNX
    from real_time_vad.engine.onnx_weights import V4_WEIGHT_FLOATS, canonical_blob_v4
    b16 = canonical_blob_v4(V4_ONNX)
    b8 = canonical_blob_v4(V4_ONNX, branch="8k")
    assert b16.size == b8.size == V4_WEIGHT_FLOATS
    assert np.isfinite(b8).all() and not np.array_equal(b16, b8)
    assert b8[-1] != b16[-1]                                # decoder bias differs between the two sub-models
    with pytest.raises(ValueError):
        canonical_blob_v4(V4_ONNX, branch="44k")
