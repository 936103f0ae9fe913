"""Measurement: cycles per tcgen05.mma with the B operand K-major SW128 vs MN-major SW64 (tools/tc_mma_rate.py's probe)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[2] / "cutter-vad_b200"))
from real_time_vad.engine import capi  # noqa: E402

L = capi.dev_lib()
out = np.zeros(2, np.int64)
print("M   N  acc  K-major  MN-major   (cycles / MMA, one CTA)")
for M, n_acc in ((128, 2), (64, 2)):
    for N in (32, 64, 96):
        r = []
        for mn in (0, 1):
            rc = L.cvad_tc_rate_mn(0, M, N, 512, 8, n_acc, 1, mn, out.ctypes.data)
            assert rc == 0, L.cvad_dev_last_error()
            r.append(out[0] / out[1])
        print(f"{M:3d} {N:3d} {n_acc:3d} {r[0]:8.1f} {r[1]:8.1f}")
