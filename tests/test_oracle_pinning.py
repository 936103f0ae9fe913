"""CPU: pin the oracle before trusting it.

Golden vectors in tests/golden/ were produced by the reference's UNMODIFIED Python
(/root/reference/src) running on the op-by-op ONNX interpreter (oracle/make_golden.py).
They pin (a) the plain-C restatement used as the GPU checker and CPU baseline,
(b) the state machine, against the expectations of the reference's own tests
(tests/test_silero_model.py:542-614, :870-954), (c) framing (tests/test_audio_utils.py:138-168)
and (d) the end-to-end known answer SampleVoiceMono.wav -> 4 segments
(websocket_service/README.md:290), and (e) the model arithmetic itself against the third-party executors that exist
in this image -- OpenCV's DNN module running the reference's own graphs (v5, v4 16 kHz, v4 8 kHz) and PyTorch's
conv1d / LSTMCell fed the file's tensors.  onnxruntime itself is absent from the image; see DESIGN.md.
"""
import numpy as np
import pytest

from conftest import GOLDEN, V5_ONNX

from vad_oracle import StateMachine, events_from_flags, sm_run_c, split_into_frames


def test_c_restatement_matches_reference_python_on_fixed_frames(ref_v5):
    g = np.load(GOLDEN / "v5_frames.npz")
    for name in ("zeros", "sine440", "noise"):
        x = g[f"in_{name}"]
        h = np.zeros(128, np.float32)
        c = np.zeros(128, np.float32)
        got = [ref_v5.frame(x, h, c)[0] for _ in range(6)]
        assert np.abs(np.array(got) - g[f"p_{name}"]).max() <= 5e-6, name
        st = g[f"state_{name}"]
        assert np.abs(h - st[0, 0]).max() <= 2e-5 and np.abs(c - st[1, 0]).max() <= 1e-4


def test_survey_probe_values_for_zero_frames():
    """SURVEY.md section 8c lists these (hand restatement + interpreter agreed to all digits)."""
    g = np.load(GOLDEN / "v5_frames.npz")
    want = [0.0442627, 0.0336124, 0.0221236, 0.0163143, 0.0121521, 0.0092669]
    assert np.abs(g["p_zeros"] - np.array(want)).max() < 2e-7


def test_onnx_interpreter_fp32_vs_fp64_bound():
    from onnx_interp import OnnxInterpreter
    g = np.load(GOLDEN / "v5_frames.npz")
    worst = 0.0
    for dt in (np.float32, np.float64):
        it = OnnxInterpreter(str(V5_ONNX), dt)
        st = np.zeros((2, 1, 128), np.float32)
        ps = []
        for _ in range(6):
            out, st = it.run({"input": g["in_noise"][None], "state": st.astype(np.float32),
                              "sr": np.array([16000], np.int64)})
            ps.append(float(out[0, 0]))
        if dt is np.float32:
            assert np.abs(np.array(ps) - g["p_noise"]).max() < 1e-7   # the golden IS this path
            p32 = np.array(ps)
        else:
            worst = np.abs(np.array(ps) - p32).max()
    assert worst < 1e-5


@pytest.mark.parametrize("mode,frame_len,hop", [("A", 480, 480), ("B", 512, 256), ("C", 512, 512)])
def test_sample_voice_known_answer_and_probabilities(ref_v5, ref_lib, mode, frame_len, hop):
    g = np.load(GOLDEN / "sample_voice.npz")
    if mode == "A":
        x = g["q16k"].astype(np.float32) / np.float32(32767.0)
        cfg = (0.4, 0.3, 0.8, 0.95, 6, 12)
    else:
        x = g["y16k"]
        cfg = (0.7, 0.7, 0.8, 0.95, 10, 50)
    want_p = g[f"{mode}_probs"]
    T = len(want_p)
    assert T == (len(x) - frame_len) // hop + 1 if mode != "A" else T == len(x) // 480
    probs, _, _ = ref_v5.run(x[None, :], T, hop=hop, frame_len=frame_len, denoise=True)
    assert np.abs(probs[0] - want_p).max() <= 2e-5
    flags, _ = sm_run_c(ref_lib, probs[0], *cfg)
    ev = [(i, 1 if k == "S" else 2) for i, k in events_from_flags(flags)]
    assert ev == [tuple(e) for e in g[f"{mode}_events"].tolist()]
    if mode == "A":
        assert [k for _, k in ev] == [1, 2] * 4          # exactly 4 voice segments


def test_state_machine_matches_reference_python_bit_exact(ref_lib):
    g = np.load(GOLDEN / "state_machine.npz")
    tags = sorted({k.rsplit("_", 1)[0] for k in g.files})
    assert len(tags) >= 18
    for tag in tags:
        probs, want, cfg = g[f"{tag}_probs"], g[f"{tag}_flags"], g[f"{tag}_cfg"]
        sp, ep, sr, er, ns, ne = cfg[0], cfg[1], cfg[2], cfg[3], int(cfg[4]), int(cfg[5])
        got_c, _ = sm_run_c(ref_lib, probs, sp, ep, sr, er, ns, ne)
        assert np.array_equal(got_c, want), tag
        sm = StateMachine(sp, ep, sr, er, ns, ne)
        got_py = np.array([sm.step(float(p)) for p in probs], np.uint8)
        assert np.array_equal(got_py, want), tag


def test_state_machine_reference_test_expectations(ref_lib):
    """tests/test_silero_model.py:870-922: 5x0.1, 5x0.8 -> START on the 3rd voiced frame (index 7);
    10x0.7 continuing; END on the 5th 0.2 (index 24)."""
    probs = np.array([0.1] * 5 + [0.8] * 5 + [0.7] * 10 + [0.2] * 5, np.float32)
    flags, _ = sm_run_c(ref_lib, probs, 0.7, 0.3, 0.8, 0.95, 3, 5)
    assert events_from_flags(flags) == [(7, "S"), (24, "E")]
    assert all(flags[i] & 4 for i in range(8, 25)) and not any(flags[i] & 4 for i in range(0, 8))
    # caps of the two deques (silero_model.py:620-628): N_s > 20 never starts, N_e > 100 never ends
    g = np.load(GOLDEN / "state_machine.npz")
    assert not (g["cap20_flags"] & 1).any()
    assert (g["cap100_flags"] & 1).any() and not (g["cap100_flags"] & 2).any()
    assert (g["edge20_flags"] & 1).any() and (g["edge20_flags"] & 2).any()


def test_framing_counts_like_the_reference():
    """audio.py:183: n = (len - frame)//hop + 1; 256..511 -> 0 frames, 512..767 -> 1, 768 -> 2,
    1..255 raises (negative dimension) -- SURVEY.md section 7 'Overlapped framing'."""
    for n, want in ((256, 0), (511, 0), (512, 1), (767, 1), (768, 2), (960000, 3749)):
        assert split_into_frames(np.zeros(n, np.float32), 512, 256).shape == (want, 512)
    with pytest.raises(ValueError):
        split_into_frames(np.zeros(255, np.float32), 512, 256)
    x = np.arange(10, dtype=np.float32)
    f = split_into_frames(x, 4, 2)
    assert f.tolist() == [[0, 1, 2, 3], [2, 3, 4, 5], [4, 5, 6, 7], [6, 7, 8, 9]]


def test_v4_is_ill_conditioned_on_band_limited_input(ref_v4, ref_v5):
    """Why the 8 kHz -> v4 parity bound is looser than 1e-4 (tests/test_gpu_resample.py): v4 feeds
    log(1 + 2^20 |STFT|) into the network, so the empty 4-8 kHz band of up-sampled audio turns
    rounding noise into signal.  Two CPU resamplers that agree to < 3e-7 on the audio -- scipy's FP32
    FFT and the exact operator rounded once -- move v4's probability by > 1e-4, but v5's by ~1e-6."""
    from scipy import signal
    from conftest import synth_streams
    n, T, n_in = 24, 9, 256
    up = signal.resample(synth_streams(n, 512 * T, seed=8), n_in * T, axis=1).astype(np.float32)
    R = signal.resample(np.eye(n_in), 512, axis=0)                       # [512][256], float64
    ya = np.zeros((n, 512 * T), np.float32)
    yb = np.zeros_like(ya)
    for s in range(n):
        for j in range(T):
            ch = up[s, j * n_in:(j + 1) * n_in]
            ya[s, j * 512:(j + 1) * 512] = signal.resample(ch, 512).astype(np.float32)
            yb[s, j * 512:(j + 1) * 512] = (R @ ch.astype(np.float64)).astype(np.float32)
    assert np.abs(ya - yb).max() < 1e-6
    d4 = np.abs(ref_v4.run(ya, T, denoise=False)[0] - ref_v4.run(yb, T, denoise=False)[0])
    d5 = np.abs(ref_v5.run(ya, T, denoise=False)[0] - ref_v5.run(yb, T, denoise=False)[0])
    assert d5.max() < 2e-5
    assert d4.max() > 1e-4 and d4.max() > 20 * d5.max()


def test_oracle_agrees_with_pytorch_modules_fed_the_files_raw_tensors(ref_v5):
    """Second opinion from third-party kernels (oracle/torch_reference.py): torch.nn.functional.conv1d and
    torch.nn.LSTMCell loaded with the tensors the .onnx carries under their torch names must give what the op-by-op
    ONNX interpretation (golden p_*) and its C restatement give -- in particular the LSTM gate re-ordering."""
    import torch
    from torch_reference import TorchV5
    from vad_oracle import v5_named_weights
    from conftest import synth_streams
    named = v5_named_weights(str(V5_ONNX))
    g = np.load(GOLDEN / "v5_frames.npz")
    t32 = TorchV5(named, torch.float32)
    for name in ("zeros", "sine440", "noise"):                  # the reference-Python goldens, six frames with state
        h = np.zeros((1, 128), np.float32)
        c = np.zeros((1, 128), np.float32)
        ps = []
        for _ in range(6):
            p, h, c = t32.frame(g[f"in_{name}"][None], h, c)
            ps.append(p[0])
        assert np.abs(np.array(ps) - g[f"p_{name}"]).max() <= 5e-6, name   # FP32 rounding (same bar as the C restatement)
        st = g[f"state_{name}"]
        assert np.abs(h[0] - st[0, 0]).max() <= 2e-5 and np.abs(c[0] - st[1, 0]).max() <= 1e-4
    # streams with state over 40 frames: C restatement (the GPU checker) vs PyTorch float32 and float64
    audio = synth_streams(12, 512 * 40, seed=77)
    want, _, _ = ref_v5.run(audio, 40, denoise=True)
    got32 = t32.run(audio, 40, denoise=True)
    got64 = TorchV5(named, torch.float64).run(audio, 40, denoise=True)
    assert np.abs(got32 - want).max() <= 2e-5
    assert np.abs(got64 - want).max() <= 2e-5
    assert want.max() > 0.9 and want.min() < 0.05             # the comparison covers the whole probability range


def test_v5_graph_executed_by_opencv_dnn_matches_the_goldens_and_the_c_restatement(ref_v5):
    """A third-party ONNX runtime on the reference's OWN graph.  onnxruntime is absent, OpenCV's DNN module is not: the
    16 kHz branch of silero_vad_v5.onnx, flattened (oracle/onnx_flatten.py: taken If branches inlined, shape arithmetic
    folded, compute nodes / attributes / weights as the file has them), is executed by cv2.dnn with the state fed
    back frame after frame, and must give the golden probabilities and states that the reference's Python produced on
    the numpy interpreter, and what the C restatement (the GPU checker) gives on speech-like streams."""
    from conftest import require_cv2
    require_cv2()
    from onnx_flatten import OpenCVSession
    from conftest import synth_streams
    g = np.load(GOLDEN / "v5_frames.npz")
    z = {"input": np.zeros((1, 512), np.float32), "state": np.zeros((2, 1, 128), np.float32), "sr": np.array([16000], np.int64)}
    sess = OpenCVSession(str(V5_ONNX), z, ["input", "state"])
    assert {"Conv", "LSTM", "Pad", "Sqrt", "Sigmoid"} <= set(sess.ops)
    for name in ("zeros", "sine440", "noise"):
        st = np.zeros((2, 1, 128), np.float32)
        ps = []
        for _ in range(6):
            out, st = sess.run({"input": g[f"in_{name}"][None], "state": st})
            ps.append(float(out.reshape(-1)[0]))
        assert np.abs(np.array(ps) - g[f"p_{name}"]).max() <= 5e-6, name
        assert np.abs(st - g[f"state_{name}"]).max() <= 1e-4
    audio = synth_streams(4, 512 * 40, seed=77)
    want, _, _ = ref_v5.run(audio, 40, denoise=True)
    for s in range(4):
        st = np.zeros((2, 1, 128), np.float32)
        for j in range(40):
            f = audio[s, j * 512:(j + 1) * 512]
            f = np.where(np.abs(f) > 0.01, f, 0.0).astype(np.float32)
            out, st = sess.run({"input": f[None], "state": st})
            # two FP32 implementations with different accumulation orders, 40 stateful frames: 2e-5 observed
            assert abs(float(out.reshape(-1)[0]) - want[s, j]) <= 5e-5, (s, j)
    assert want.max() > 0.9 and want.min() < 0.05


@pytest.mark.parametrize("sr", [16000, 8000])
def test_v4_graph_executed_by_opencv_dnn_matches_the_interpreter(sr):
    """Same pin for silero_vad.onnx: its 16 kHz branch and its 8 kHz sub-model (two LSTM time steps per frame)."""
    from conftest import require_cv2
    require_cv2()
    from onnx_flatten import OpenCVSession
    from onnx_interp import OnnxInterpreter
    from conftest import V4_ONNX, synth_streams
    z = {"input": np.zeros((1, 512), np.float32), "h": np.zeros((2, 1, 64), np.float32), "c": np.zeros((2, 1, 64), np.float32),
         "sr": np.array([sr], np.int64)}
    sess = OpenCVSession(str(V4_ONNX), z, ["input", "h", "c"])
    assert sess.ops.count("LSTM") == 2 and "Log" in sess.ops
    it = OnnxInterpreter(str(V4_ONNX), np.float32)
    audio = synth_streams(2, 512 * 12, seed=5)[1] + 0.02 * np.random.default_rng(1).standard_normal(512 * 12).astype(np.float32)
    h = c = np.zeros((2, 1, 64), np.float32)
    h2, c2 = h.copy(), c.copy()
    worst = 0.0
    for j in range(12):
        f = audio[j * 512:(j + 1) * 512].astype(np.float32)[None]
        out, h, c = sess.run({"input": f, "h": h, "c": c})
        ref, h2, c2 = it.run({"input": f, "h": h2, "c": c2, "sr": np.array([sr], np.int64)})
        worst = max(worst, abs(float(out.reshape(-1)[0]) - float(np.asarray(ref).reshape(-1)[0])))
    assert worst <= 2e-5, worst
