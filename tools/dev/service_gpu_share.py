"""How much of BatchedVADManager.step()'s GPU phase is kernel time?  (10,000 live streams, bench.service_tick's recipe, with
cvad_set_timing on the manager's engine.)"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "cutter-vad_b200"))
import bench  # noqa: E402
from real_time_vad import BatchedVADManager  # noqa: E402

orig = BatchedVADManager.__init__
box = {}


def spy(self, *a, **k):
    orig(self, *a, **k)
    box["mgr"] = self
    self._engine.set_timing(True)


BatchedVADManager.__init__ = spy
close = BatchedVADManager.close


def spy_close(self):
    fe, rec, n = self._engine.read_timing()
    print(f"kernels: {n} steps timed, {1e3 * (fe + rec) / max(n, 1):.1f} us per step (front end {1e3 * fe / max(n, 1):.1f}, recurrent {1e3 * rec / max(n, 1):.1f})")
    close(self)


BatchedVADManager.close = spy_close
r = bench.service_tick(10000, ticks=60)
print({k: v for k, v in r.items() if "native" in k or "step_ms" in k})
