"""GPU: the FP16-split build of the v5 kernels (CVAD_MATH_TC16; fused one-frame kernel and the two-kernel multi-frame
form) -- two FP16 operand parts, three tensor-core products per MAC, every activation operand scaled per STREAM by a
power of two.  Besides the oracle bar
(1e-4, also run over the whole v5 parity suite through its math_mode fixture) this file checks what the scaling must
guarantee: a stream's result does not depend on its neighbours in the batch, and loud / quiet / out-of-range streams
in one tile stay on the oracle."""
import numpy as np
import pytest

from conftest import synth_streams

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _steps(eng, audio, T, **kw):
    return np.stack([eng.step(np.ascontiguousarray(audio[:, j * 512:(j + 1) * 512]), **kw).probs[:, 0] for j in range(T)], axis=1)


def test_tc16_matches_oracle_and_bf16_split(engine_factory, ref_v5):
    n, T = 300, 40
    audio = synth_streams(n, 512 * T, seed=61)
    want, _, _ = ref_v5.run(audio, T, denoise=True)
    eng = engine_factory(512, math="tc16")
    assert eng.math == "tc16"
    eng.reset()
    eng.configure(enable_denoising=True)
    got = _steps(eng, audio, T)
    assert np.abs(got - want).max() <= TOL
    ref = engine_factory(512, math="tc")
    ref.reset()
    ref.configure(enable_denoising=True)
    other = _steps(ref, audio, T)
    assert np.abs(got - other).max() <= 3e-5
    # a multi-frame call under tc16 runs the FP16-split two-kernel path: on the oracle, next to the BF16 split, and
    # next to its own one-frame form
    eng.reset()
    ref.reset()
    many = eng.step(audio).probs
    assert np.abs(many - want).max() <= TOL
    assert np.abs(many - ref.step(audio).probs).max() <= 3e-5
    assert np.abs(many - got).max() <= 3e-5


def test_tc16_result_does_not_depend_on_the_neighbours(engine_factory):
    """Per-stream (not per-tile) scaling: the same stream gives the same bits alone, among quiet streams, among
    streams 10^6 times louder, and at any position of the batch."""
    T = 12
    probe = synth_streams(3, 512 * T, seed=5)
    rng = np.random.default_rng(1)
    eng = engine_factory(512, math="tc16")
    eng.configure(enable_denoising=False)
    eng.reset()
    alone = _steps(eng, probe, T)
    for scale, where in ((1e-5, 0), (1.0, 17), (1e6, 40), (3e4, 93)):
        crowd = (scale * rng.standard_normal((96, 512 * T))).astype(np.float32)
        crowd[where:where + 3] = probe
        eng.reset()
        got = _steps(eng, crowd, T)
        assert np.array_equal(got[where:where + 3], alone), (scale, where)
    # the same property for the multi-frame (two-kernel) form
    eng.reset()
    alone_many = eng.step(probe).probs
    crowd = (1e5 * rng.standard_normal((96, 512 * T))).astype(np.float32)
    crowd[50:53] = probe
    eng.reset()
    assert np.array_equal(eng.step(crowd).probs[50:53], alone_many)


def test_tc16_mixed_amplitudes_in_one_tile_stay_on_the_oracle(engine_factory, ref_v5):
    T = 16
    base = synth_streams(32, 512 * T, seed=13)
    gains = np.array([1e-6, 1e-4, 1e-2, 0.1, 1, 3, 30, 3e3, 32768, 1e6, 1e8, 0] * 3, np.float32)[:32]
    audio = (base * gains[:, None]).astype(np.float32)
    want, _, _ = ref_v5.run(audio, T, denoise=False)
    eng = engine_factory(512, math="tc16")
    eng.reset()
    eng.configure(enable_denoising=False)
    got = _steps(eng, audio, T)
    assert np.isfinite(got).all()
    assert np.abs(got - want).max() <= TOL, np.abs(got - want).max(1)


def test_tc16_int16_short_frames_slots_and_idle_streams(engine_factory, ref_v5):
    from real_time_vad.engine import capi
    n, T = 45, 10
    audio = synth_streams(n, 480 * T, seed=23)
    q = np.clip(np.round(audio * 32767.0), -32768, 32767).astype(np.int16)
    xf = q.astype(np.float32) / np.float32(32767.0)
    want, _, _ = ref_v5.run(xf, T, hop=480, frame_len=480, denoise=True)
    eng = engine_factory(512, math="tc16")
    eng.reset()
    eng.configure(enable_denoising=True)
    slots = (np.arange(n) * 7 + 3).astype(np.int32)
    got = np.zeros((n, T), np.float32)
    idle = np.arange(n) % 5 == 0
    done = np.zeros(n, int)
    for j in range(2 * T):                       # streams of the idle group only run on every other step
        run = np.where(idle & (j % 2 == 1), 0, 1).astype(np.int32)
        run[done >= T] = 0
        block = np.stack([q[s, done[s] * 480:done[s] * 480 + 480] if done[s] < T else np.zeros(480, np.int16) for s in range(n)])
        r = eng.step(block, slots=slots, n_frames=run, max_frames=1, frame_len=480, hop=480, pcm_format=capi.PCM_S16_32767)
        for s in np.flatnonzero(run):
            got[s, done[s]] = r.probs[s, 0]
        done += run
    assert (done == T).all()
    assert np.abs(got - want).max() <= TOL


@pytest.mark.parametrize("math", ["tc16", "tc", "fp32"])
def test_engine_against_the_pytorch_reference(engine_factory, math):
    """The CUDA path against third-party kernels directly (oracle/torch_reference.py: torch conv1d + LSTMCell fed the
    .onnx file's tensors), float32, 1e-4: one frame per step and all frames in one call."""
    import torch
    from conftest import V5_ONNX
    from torch_reference import TorchV5
    from vad_oracle import v5_named_weights
    n, T = 64, 30
    audio = synth_streams(n, 512 * T, seed=88)
    want = TorchV5(v5_named_weights(str(V5_ONNX)), torch.float32).run(audio, T, denoise=True)
    eng = engine_factory(512, math=math)
    eng.reset()
    eng.configure(enable_denoising=True)
    got = _steps(eng, audio, T)
    assert np.abs(got - want).max() <= TOL
    eng.reset()
    assert np.abs(eng.step(audio).probs - want).max() <= TOL


@pytest.mark.parametrize("math", ["tc16", "fp32"])
def test_engine_against_opencv_dnn_on_the_reference_graph(engine_factory, math):
    """The CUDA path against a third-party ONNX runtime executing the reference's own graph (oracle/onnx_flatten.py:
    silero_vad_v5.onnx's 16 kHz branch inside OpenCV's DNN module, state fed back per frame), 1e-4."""
    from conftest import require_cv2
    require_cv2()
    from conftest import V5_ONNX
    from onnx_flatten import OpenCVSession
    n, T = 6, 30
    audio = synth_streams(n, 512 * T, seed=19)
    z = {"input": np.zeros((1, 512), np.float32), "state": np.zeros((2, 1, 128), np.float32), "sr": np.array([16000], np.int64)}
    sess = OpenCVSession(str(V5_ONNX), z, ["input", "state"])
    want = np.zeros((n, T))
    for s in range(n):
        st = np.zeros((2, 1, 128), np.float32)
        for j in range(T):
            f = audio[s, j * 512:(j + 1) * 512]
            out, st = sess.run({"input": np.where(np.abs(f) > 0.01, f, 0.0).astype(np.float32)[None], "state": st})
            want[s, j] = float(out.reshape(-1)[0])
    eng = engine_factory(512, math=math)
    eng.reset()
    eng.configure(enable_denoising=True)
    got = _steps(eng, audio, T)
    assert np.abs(got - want).max() <= TOL
    assert want.max() > 0.8 and want.min() < 0.1
