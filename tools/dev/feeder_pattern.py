"""The feeder's call pattern on the bare engine: a one-frame step over a RANDOM 87 % subset of 10,000 slots (pinned int16
frames of 480 samples), alone and with a second small step submitted behind it before the first is collected."""
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


def pinned(shape):
    nb = int(np.prod(shape)) * 2
    ptr = L.cvad_alloc_pinned(nb)
    return np.frombuffer((C.c_char * nb).from_address(ptr), np.int16).reshape(shape)


kw = dict(frame_len=480, hop=480, pcm_format=capi.PCM_S16_32767)
big = pinned((n, 480))
big[:] = (rng.standard_normal((n, 480)) * 3000).astype(np.int16)
small = pinned((1300, 480))
small[:] = big[:1300]
sub = np.sort(rng.choice(n, 8750, replace=False)).astype(np.int32)
sub2 = sub[:1300].copy()
cases = {
    "8750 contiguous slots": lambda: eng.step(big[:8750], slots=np.arange(8750, dtype=np.int32), **kw),
    "8750 random slots": lambda: eng.step(big[:8750], slots=sub, **kw),
    "8750 random + 1300 behind it (submit, submit, collect, collect)": None,
}
for label, fn in cases.items():
    ts = []
    for _ in range(60):
        t0 = time.perf_counter()
        if fn is not None:
            fn()
        else:
            a = eng.submit(big[:8750], slots=sub, **kw)
            b = eng.submit(small, slots=sub2, **kw)
            a.collect()
            b.collect()
        ts.append(time.perf_counter() - t0)
    print(f"{label:64s} wall p50 {1e3 * np.median(ts[10:]):.3f} ms")
# the same with the GPU idle between steps, as it is in a service that does ~2 ms of host work per tick
for idle_ms in (0.5, 2.0, 10.0):
    ts = []
    for _ in range(40):
        t_end = time.perf_counter() + idle_ms * 1e-3
        while time.perf_counter() < t_end:
            pass
        t0 = time.perf_counter()
        eng.step(big[:8750], slots=sub, **kw)
        ts.append(time.perf_counter() - t0)
    print(f"8750 random slots after {idle_ms:4.1f} ms of idle GPU (host busy)               wall p50 {1e3 * np.median(ts[5:]):.3f} ms")
eng.set_timing(True)
for _ in range(20):
    eng.step(big[:8750], slots=sub, **kw)
fe, rec, k = eng.read_timing()
print(f"kernels, random slots: {1e3 * (fe + rec) / k:.1f} us per step")
for _ in range(20):
    eng.step(big[:8750], slots=np.arange(8750, dtype=np.int32), **kw)
fe, rec, k = eng.read_timing()
print(f"kernels, contiguous slots: {1e3 * (fe + rec) / k:.1f} us per step")
eng.close()
