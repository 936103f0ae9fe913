"""Is a host-to-device copy of FRESHLY WRITTEN pinned memory slower than of memory that has left the CPU caches?"""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "cutter-vad_b200"))
from real_time_vad.engine import capi  # noqa: E402
from real_time_vad.engine.stream_engine import StreamEngine  # noqa: E402

n = 10000
eng = StreamEngine(max_streams=n)
L = capi.lib()
rng = np.random.default_rng(0)
a16 = (rng.standard_normal((n, 480)) * 3000).astype(np.int16)
ptr = L.cvad_alloc_pinned(a16.nbytes)
pin = np.frombuffer((C.c_char * a16.nbytes).from_address(ptr), np.int16).reshape(a16.shape)
pin[:] = a16
kw = dict(frame_len=480, hop=480, pcm_format=capi.PCM_S16_32767)
for _ in range(10):
    eng.step(pin, **kw)
for label, rewrite in (("stale (written once)", False), ("rewritten before every step", True), ("stale again", False)):
    ts = []
    for _ in range(100):
        if rewrite:
            pin[:] = a16
        t0 = time.perf_counter()
        eng.step(pin, **kw)
        ts.append(time.perf_counter() - t0)
    print(f"{label:32s} step wall p50 {1e3 * np.median(ts):.3f} ms")
eng.close()
