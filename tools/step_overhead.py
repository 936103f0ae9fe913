#!/usr/bin/env python
"""Where a host-buffer cvad_step spends its wall time: service-mode shape (10,000 streams x one 480-sample int16 frame,
slot list + ragged frame counts) and the bench shape (4,096 x 512 float32), pinned and pageable input."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT / "cutter-vad_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
from real_time_vad.engine import capi  # noqa: E402
from real_time_vad.engine.stream_engine import StreamEngine  # noqa: E402


def run(label, eng, audio, reps=200, **kw):
    for _ in range(10):
        eng.step(audio, **kw)
    eng.set_timing(True)
    t0 = time.perf_counter()
    for _ in range(reps):
        eng.step(audio, **kw)
    wall = (time.perf_counter() - t0) / reps
    fe, rec, k = eng.read_timing()
    eng.set_timing(False)
    print(f"{label:46s} wall {1e3 * wall:.3f} ms   kernels {(fe + rec) / k:.3f} ms   bytes in {audio.nbytes / 1e6:.1f} MB")


def main():
    n = 10000
    eng = StreamEngine(max_streams=n)
    rng = np.random.default_rng(0)
    a16 = (rng.standard_normal((n, 480)) * 3000).astype(np.int16)
    slots = np.arange(n, dtype=np.int32)
    nfr = np.ones(n, np.int32)
    kw = dict(frame_len=480, hop=480, pcm_format=capi.PCM_S16_32767)
    run("10000 x 480 s16, pageable, slots + n_frames", eng, a16, slots=slots, n_frames=nfr, max_frames=1, **kw)
    run("10000 x 480 s16, pageable, plain", eng, a16, **kw)
    L = capi.lib()
    ptr = L.cvad_alloc_pinned(a16.nbytes)
    import ctypes as C
    pin = np.frombuffer((C.c_char * a16.nbytes).from_address(ptr), np.int16).reshape(a16.shape)
    pin[:] = a16
    run("10000 x 480 s16, pinned, slots + n_frames", eng, pin, slots=slots, n_frames=nfr, max_frames=1, **kw)
    run("10000 x 480 s16, pinned, plain", eng, pin, **kw)
    f32 = (0.1 * rng.standard_normal((4096, 512))).astype(np.float32)
    run("4096 x 512 f32, pageable, plain", eng, f32)
    eng.close()


if __name__ == "__main__":
    main()
