"""ORACLE (test infrastructure): generate tests/golden/*.npz by running the REFERENCE's own,
unmodified Python (/root/reference/src) on top of the onnxruntime-shaped numpy interpreter
(oracle/ort_shim).  Run in the build container only -- /root/reference does not exist on
the GPU box, which is why the outputs are committed:

    python oracle/make_golden.py

Vectors:
  v5_frames.npz        6 repeated frames (zeros / 440 Hz sine) through SileroVADModel.predict
  sample_voice.npz     examples/audios/SampleVoiceMono.wav resampled with AudioUtils.resample_audio,
                       through VADWrapper in three framings:
                         A  websocket-like: 480-sample int16-quantised messages, 0.4/0.3/6/12
                            (websocket_service/server/vad_websocket_server.py:341, :565-572)
                         B  one process_audio_data call, VADConfig defaults (hop 256)
                         C  512 samples per call, VADConfig defaults
                       per-frame probabilities, event (frame, kind) lists, WAV byte counts + sha256
  sample_voice_v4.npz  the same file through VADWrapper(model_version=V4): framings A and B (events, probabilities, WAV hashes)
  sample_voice_stereo.npz  examples/audios/SampleVoiceStereo.wav as a 2-D [samples, 2] array through VADWrapper
                       (AudioUtils.convert_to_mono, audio.py:193-208).  Both channels of that file equal
                       SampleVoiceMono.wav sample for sample (asserted here, sha256 of the PCM recorded), so the tests
                       rebuild the stereo input from the committed mono file; a second case with DIFFERENT channels
                       (right = left delayed and attenuated) pins the float32 mean itself.
  state_machine.npz    scripted probability sequences through VADProcessor._process_voice_state
                       (the cases of tests/test_silero_model.py:542-614, :870-954 plus random ones)
"""
import hashlib
import os
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference")
sys.path[:0] = [str(HERE / "ort_shim"), str(HERE), str(REF / "src")]
OUT = HERE.parent / "tests" / "golden"


def main():
    from scipy.io import wavfile
    import real_time_vad
    assert str(REF) in real_time_vad.__file__, real_time_vad.__file__
    from real_time_vad import VADConfig, VADWrapper, SampleRate, SileroModelVersion
    from real_time_vad.core.silero_model import SileroVADModel, VADProcessor
    from real_time_vad.utils.audio import AudioUtils

    models = REF / "src" / "real_time_vad" / "models"

    # ---------------------------------------------------------------- v5_frames
    m = SileroVADModel(str(models / "silero_vad_v5.onnx"), SileroModelVersion.V5)
    frames = {"zeros": np.zeros(512, np.float32),
              "sine440": (0.5 * np.sin(2 * np.pi * 440 * np.arange(512) / 16000)).astype(np.float32)}
    rng = np.random.default_rng(42)
    frames["noise"] = (0.1 * rng.standard_normal(512)).astype(np.float32)
    out = {}
    for name, x in frames.items():
        m.reset()
        out[f"in_{name}"] = x
        out[f"p_{name}"] = np.array([m.predict(x, 16000) for _ in range(6)], np.float64)
        out[f"state_{name}"] = m.model_state.state.copy()
    np.savez_compressed(OUT / "v5_frames.npz", **out)

    # ---------------------------------------------------------------- sample_voice
    sr, x = wavfile.read(str(REF / "examples" / "audios" / "SampleVoiceMono.wav"))
    y = AudioUtils.resample_audio(x.astype(np.float32) / 32768.0, sr, 16000)
    res = {"y16k": y}

    def run(tag, cfg, chunks):
        w = VADWrapper(cfg)
        probs, events, wavs = [], [], []
        orig = w.processor.model.predict

        def spy(chunk, rate):
            p = orig(chunk, rate)
            probs.append(p)
            return p
        object.__setattr__(w.processor.model, "predict", spy)
        w.set_callbacks(voice_start_callback=lambda: events.append((len(probs) - 1, 1)),
                        voice_end_callback=lambda b: (events.append((len(probs) - 1, 2)), wavs.append(b)))
        for c in chunks:
            w.process_audio_data(c)
        res[f"{tag}_probs"] = np.array(probs, np.float64)
        res[f"{tag}_events"] = np.array(events, np.int64).reshape(-1, 2)
        res[f"{tag}_wav_len"] = np.array([len(b) for b in wavs], np.int64)
        res[f"{tag}_wav_sha"] = np.array([hashlib.sha256(b).hexdigest() for b in wavs])
        print(tag, len(probs), "frames", events)

    def run_into(store, tag, cfg, chunks):
        saved = dict(res)
        res.clear()
        run(tag, cfg, chunks)
        store.update(res)
        res.clear()
        res.update(saved)

    q = np.clip(np.round(y * 32767.0), -32768, 32767).astype(np.int16)
    res["q16k"] = q
    cfg_a = VADConfig(sample_rate=SampleRate.SAMPLERATE_16, buffer_size=480, vad_start_probability=0.4,
                      vad_end_probability=0.3, voice_start_frame_count=6, voice_end_frame_count=12,
                      voice_start_ratio=0.8, voice_end_ratio=0.95)
    run("A", cfg_a, [q[i * 480:(i + 1) * 480].astype(np.float32) / 32767.0 for i in range(len(q) // 480)])
    run("B", VADConfig(), [y])
    run("C", VADConfig(), [y[i * 512:(i + 1) * 512] for i in range(len(y) // 512)])
    np.savez_compressed(OUT / "sample_voice.npz", **res)

    # ---------------------------------------------------------------- sample_voice_v4 (model_version=V4 through the reference)
    v4 = {}
    cfg_a4 = cfg_a.model_copy(update={"model_version": SileroModelVersion.V4})
    run_into(v4, "A", cfg_a4, [q[i * 480:(i + 1) * 480].astype(np.float32) / 32767.0 for i in range(len(q) // 480)])
    run_into(v4, "B", VADConfig(model_version=SileroModelVersion.V4), [y])
    np.savez_compressed(OUT / "sample_voice_v4.npz", **v4)

    # ---------------------------------------------------------------- sample_voice_stereo
    sr2, xs = wavfile.read(str(REF / "examples" / "audios" / "SampleVoiceStereo.wav"))
    assert sr2 == sr and xs.shape == (len(x), 2) and np.array_equal(xs[:, 0], x) and np.array_equal(xs[:, 1], x)
    st = {"stereo_pcm_sha256": np.array(hashlib.sha256(np.ascontiguousarray(xs).tobytes()).hexdigest())}
    ys = np.stack([y, y], axis=1)                                      # [samples, 2] at 16 kHz, like the file's two channels
    run_into(st, "file", VADConfig(), [ys])
    # channels that differ: right = left delayed by 37 samples and attenuated -> np.mean(axis=1) in float32
    right = np.concatenate([np.zeros(37, np.float32), y[:-37]]) * np.float32(0.61)
    yd = np.stack([y, right], axis=1).astype(np.float32)
    st["diff_right"] = right
    run_into(st, "diff", VADConfig(vad_start_probability=0.4, vad_end_probability=0.3, voice_start_frame_count=6,
                                   voice_end_frame_count=12), [yd])
    np.savez_compressed(OUT / "sample_voice_stereo.npz", **st)

    # ---------------------------------------------------------------- state_machine
    cases = {}
    rng = np.random.default_rng(7)

    def sm_case(tag, probs, **cfgkw):
        cfg = VADConfig(**cfgkw)
        p = VADProcessor(cfg)
        fl = []
        frame = np.full(512, 0.5, np.float32)
        for pr in probs:
            r = p._process_voice_state(float(np.float32(pr)), frame)
            fl.append(int(r["voice_started"]) | (int(r["voice_ended"]) << 1) | (int(r["voice_continuing"]) << 2))
        cases[f"{tag}_probs"] = np.array(probs, np.float32)
        cases[f"{tag}_flags"] = np.array(fl, np.uint8)
        cases[f"{tag}_cfg"] = np.array([cfg.vad_start_probability, cfg.vad_end_probability, cfg.voice_start_ratio,
                                        cfg.voice_end_ratio, cfg.voice_start_frame_count,
                                        cfg.voice_end_frame_count], np.float64)

    kw = dict(vad_start_probability=0.7, vad_end_probability=0.3, voice_start_frame_count=3, voice_end_frame_count=5)
    sm_case("cycle", [0.1] * 5 + [0.8] * 5 + [0.7] * 10 + [0.2] * 5, **kw)          # test_silero_model.py:870-922
    sm_case("three", ([0.1] * 3 + [0.8] * 4 + [0.2] * 6) * 3, **kw)                  # :924-954
    sm_case("ties", [0.7, 0.7, 0.7, 0.3, 0.3, 0.29999998, 0.3, 0.2, 0.2, 0.2, 0.2, 0.2, 0.7], **kw)
    sm_case("cap20", [0.9] * 60, vad_start_probability=0.5, vad_end_probability=0.5,
            voice_start_frame_count=21, voice_end_frame_count=5)                      # N_s > 20 can never start
    sm_case("cap100", [0.9] * 5 + [0.1] * 150, vad_start_probability=0.5, vad_end_probability=0.5,
            voice_start_frame_count=2, voice_end_frame_count=101)                     # N_e > 100 can never end
    sm_case("edge20", [0.9] * 25 + [0.1] * 110, vad_start_probability=0.5, vad_end_probability=0.5,
            voice_start_frame_count=20, voice_end_frame_count=100)
    for i in range(12):
        n = 400
        base = np.clip(0.5 + 0.45 * np.sin(np.arange(n) / rng.uniform(3, 30)) + 0.2 * rng.standard_normal(n), 0, 1)
        sm_case(f"rand{i}", base.astype(np.float32),
                vad_start_probability=float(rng.choice([0.4, 0.5, 0.7])),
                vad_end_probability=float(rng.choice([0.3, 0.35, 0.7])),
                voice_start_frame_count=int(rng.integers(1, 12)), voice_end_frame_count=int(rng.integers(1, 30)),
                voice_start_ratio=float(rng.choice([0.5, 0.8, 1.0])), voice_end_ratio=float(rng.choice([0.6, 0.95, 1.0])))
    np.savez_compressed(OUT / "state_machine.npz", **cases)
    print("written to", OUT)


if __name__ == "__main__":
    main()
