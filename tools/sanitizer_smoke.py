"""Small workload for compute-sanitizer: every tensor-core kernel once (fused one-frame step, two-kernel multi-frame
step, mixed-rate resampling, v4 STFT), checked against nothing -- the tool's report is the result."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT / "cutter-vad_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
from conftest import synth_streams  # noqa: E402
from real_time_vad.engine.stream_engine import StreamEngine  # noqa: E402

x = synth_streams(70, 512 * 3, seed=1)
eng = StreamEngine(max_streams=128, math="tc")
eng.step(x[:, :512])                       # fused kernel
eng.step(x)                                # front end + recurrent kernels
rates = np.array([48000, 24000, 16000, 8000] * 18, np.int32)[:70]
big = np.zeros((70, 1536 * 2), np.float32)
big[:, :1024] = x[:, :1024]
eng.step(big, src_rates=rates, max_frames=2)
eng.close()
# the default build: FP16-split fused kernel (MN-major operands, direct state rows, gate-pipelined cell), identity slots and
# a slot list with gaps (the gather path of the state), chained device-pointer steps left to tests/test_gpu_device_steps.py
e16 = StreamEngine(max_streams=128, math="tc16")
e16.step(x[:, :512])
e16.step(x[:, 512:1024])
e16.step(x[:37, :512], slots=np.arange(0, 74, 2, dtype=np.int32))
e16.step(x)
e16.close()
e4 = StreamEngine("v4", max_streams=64, math="tc")
e4.step(x[:40])
e4.close()
print("sanitizer smoke done")
