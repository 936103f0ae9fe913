"""e2e leg of bench.py (int16 wire format) with 2 / 3 / 4 steps in flight."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "cutter-vad_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
from real_time_vad.engine.stream_engine import StreamEngine  # noqa: E402

torch.cuda.set_device(0)
wl = bench.Workload(torch, 0, 0, "v5", 4096, 1, 16000, False)
eng = StreamEngine(max_streams=4096, math="tc16")
for depth in (2, 3, 4, 3, 4):
    r = bench.time_e2e(torch, None, 1, eng, wl, 200, 20, "s16", depth=depth)
    print(depth, {k: (round(v, 1) if isinstance(v, float) else v) for k, v in r.items() if k in ("value", "ms_per_step", "value_max")})
