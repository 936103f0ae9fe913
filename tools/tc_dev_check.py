"""Development check of the tensor-core path: per-layer dump and a few recurrent steps vs the oracle."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT / "cutter-vad_b200", ROOT / "oracle", ROOT / "tests"):
    sys.path.insert(0, str(p))
from conftest import V5_ONNX, synth_streams  # noqa: E402
from real_time_vad.engine.stream_engine import StreamEngine  # noqa: E402
from vad_oracle import RefLib, RefV5, v5_blob  # noqa: E402

ref = RefV5(v5_blob(str(V5_ONNX)), RefLib())
eng = StreamEngine(max_streams=128, math="tc")
eng.configure(enable_denoising=False)
x = synth_streams(32, 16000 + 512, seed=3)[:, 16000:16000 + 512].copy()
got = eng.debug_dump(x)
n = 32
want = {"mag": np.zeros((129, 3, n), np.float32), "e0": np.zeros((128, 3, n), np.float32),
        "e1": np.zeros((64, 2, n), np.float32), "e2": np.zeros((64, n), np.float32), "feat": np.zeros((128, n), np.float32)}
for s in range(n):
    _, d = ref.frame(x[s], np.zeros(128, np.float32), np.zeros(128, np.float32), want_dbg=True)
    want["mag"][:, :, s] = d[0:387].reshape(129, 3)
    want["e0"][:, :, s] = d[387:771].reshape(128, 3)
    want["e1"][:, :, s] = d[771:899].reshape(64, 2)
    want["e2"][:, s] = d[899:963]
    want["feat"][:, s] = d[963:1091]
for name in ("mag", "e0", "e1", "e2", "feat"):
    w = want[name]
    g = got[name][..., :w.shape[-1]]
    err = np.abs(g - w)
    print(f"{name:5s} max err {err.max():.3e} (scale {np.abs(w).max():.3f}) worst at {np.unravel_index(err.argmax(), err.shape)}")

for n_streams in (1, 32, 70):
    eng.reset()
    eng.configure(enable_denoising=True)
    n_steps = 12
    audio = synth_streams(n_streams, 512 * n_steps, seed=5)
    wantp, h_ref, c_ref = ref.run(audio, n_steps, hop=512, frame_len=512, denoise=True)
    gotp = np.zeros_like(wantp)
    for j in range(n_steps):
        r = eng.step(audio[:, j * 512:(j + 1) * 512])
        gotp[:, j] = r.probs[:, 0]
    print(f"n={n_streams}: 1 frame/step max |dp| = {np.abs(gotp - wantp).max():.3e}")
    h, c, sm, fd = eng.get_state(n_streams - 1)
    print("   h err", np.abs(h - h_ref[n_streams - 1]).max(), "c err", np.abs(c - c_ref[n_streams - 1]).max(), "frames", fd)
n, T = 45, 37
eng.reset()
audio = synth_streams(n, 256 * (T - 1) + 512, seed=7)
wantp, _, _ = ref.run(audio, T, hop=256, frame_len=512, denoise=True)
r = eng.step(audio, hop=256)
print(f"hop256 T={T}: max |dp| = {np.abs(r.probs - wantp).max():.3e}")
