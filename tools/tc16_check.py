#!/usr/bin/env python
"""Development check of the FP16-split fused kernel (CVAD_MATH_TC16): one-frame steps of N streams x T frames against
the oracle and against the BF16-split build, plus the device time of a 4,096-stream step in both builds."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT / "cutter-vad_b200", ROOT / "oracle", ROOT / "tests", ROOT):
    sys.path.insert(0, str(p))


def main():
    from conftest import synth_streams, V5_ONNX
    from real_time_vad.engine.stream_engine import StreamEngine
    from vad_oracle import RefLib, RefV5, v5_blob
    n, T = int(sys.argv[1]) if len(sys.argv) > 1 else 200, 30
    audio = synth_streams(n, 512 * T, seed=7)
    audio[3] *= 1e-3
    audio[4] *= 30000.0
    audio[5] = 0.0
    ref = RefV5(v5_blob(str(V5_ONNX)), RefLib())
    want, _, _ = ref.run(audio, T, denoise=True)
    out = {}
    for math in ("tc", "tc16"):
        eng = StreamEngine("v5", max_streams=n, device=0, math=math)
        eng.configure(enable_denoising=True)
        got = np.stack([eng.step(audio[:, j * 512:(j + 1) * 512]).probs[:, 0] for j in range(T)], axis=1)
        out[math] = got
        d = np.abs(got - want)
        print(f"{math:5s} max|dp| vs oracle {d.max():.3e}  per-frame-0 {np.abs(got[:, 0] - want[:, 0]).max():.3e}  "
              f"worst stream {int(d.max(1).argmax())}  finite {np.isfinite(got).all()}")
        eng.close()
    print("tc16 vs tc", np.abs(out["tc16"] - out["tc"]).max())
    # timing
    N = 4096
    big = synth_streams(64, 512, seed=3)
    big = np.tile(big, (N // 64, 1))
    for math in ("tc", "tc16"):
        eng = StreamEngine("v5", max_streams=N, device=0, math=math)
        for _ in range(5):
            eng.step(big)
        eng.set_timing(True)
        for _ in range(50):
            eng.step(big)
        fe, rec, k = eng.read_timing()
        print(f"{math:5s} kernel {1e3 * fe / k:.1f} us per 4096-stream step")
        eng.close()


if __name__ == "__main__":
    main()
