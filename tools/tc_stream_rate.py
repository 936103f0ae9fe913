"""Measurement: bytes per SM cycle that ONE SM pulls from L2 into shared memory with cp.async.bulk (16 KB tiles,
ring of `depth` slots, nothing consuming them), alone and with 128 / 148 CTAs doing the same.  This is the supply
side of the weight-streaming kernels."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "cutter-vad_b200"))
from real_time_vad.engine import capi  # noqa: E402

L = capi.dev_lib()
out = np.zeros(2, np.int64)
print("grid depth tile   cycles/tile   B/cycle/SM")
for grid in (1, 8, 128, 148):
    for depth, tile in ((2, 16384), (4, 16384), (7, 16384), (12, 16384), (7, 8192), (4, 32768)):
        rc = L.cvad_bulk_rate(0, 2000, depth, tile, grid, 1 << 20, out.ctypes.data)
        assert rc == 0, L.cvad_dev_last_error()
        cyc = out[0] / out[1]
        print(f"{grid:4d} {depth:5d} {tile:6d} {cyc:10.1f} {tile / cyc:10.1f}")
