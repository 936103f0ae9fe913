"""Measurement: cycles per M x N x 16 BF16 tcgen05.mma with both operands in shared memory (one CTA and 148 CTAs).
Decides the tile shape of the tensor-core path."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "cutter-vad_b200"))
from real_time_vad.engine import capi  # noqa: E402

L = capi.dev_lib()
out = np.zeros(2, np.int64)
print("M   N   grid cycles/MMA  MAC/cyc  smemB/cyc")
for grid, M, n_acc in ((1, 128, 1), (1, 128, 2), (1, 128, 4), (1, 64, 1), (1, 64, 4), (148, 128, 2)):
    for N in (16, 32, 64, 96, 128, 192, 256):
        if n_acc * N > 512:
            continue
        rc = L.cvad_tc_rate(0, M, N, 512, 8, n_acc, grid, out.ctypes.data)
        assert rc == 0, L.cvad_dev_last_error()
        cyc = out[0] / out[1]
        print(f"{M:3d} {N:3d} {grid:4d} acc={n_acc} {cyc:8.1f} {M * N * 16 / cyc:8.0f} {(M + N) * 32 / cyc:8.1f}")
