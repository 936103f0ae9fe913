"""dev: 256 distinct streams x 2000 one-frame steps through each arithmetic build; probabilities to gpurun_out/ for offline analysis."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
for p in (ROOT / "cutter-vad_b200", ROOT / "tests", ROOT / "oracle"):
    sys.path.insert(0, str(p))
from conftest import synth_streams  # noqa: E402
from real_time_vad.engine.stream_engine import StreamEngine  # noqa: E402
distinct, T = 256, 2000
base = synth_streams(distinct, 512 * T, seed=123)
out = {}
for math in ("tc16", "tc", "fp32"):
    eng = StreamEngine("v5", max_streams=distinct, math=math)
    eng.configure(enable_denoising=True)
    got = np.zeros((distinct, T), np.float32)
    for j in range(T):
        got[:, j] = eng.step(base[:, j * 512:(j + 1) * 512]).probs[:, 0]
    out[math] = got
    eng.close()
    # the multi-frame (two-kernel) form of the same build, 100 frames per call
    eng = StreamEngine("v5", max_streams=distinct, math=math)
    eng.configure(enable_denoising=True)
    got2 = np.concatenate([eng.step(base[:, k * 512 * 100:(k + 1) * 512 * 100]).probs for k in range(T // 100)], axis=1)
    out[math + "_multi"] = got2
    eng.close()
np.savez_compressed(ROOT / "gpurun_out" / "drift_dump.npz", **out)
print({k: v.shape for k, v in out.items()})
