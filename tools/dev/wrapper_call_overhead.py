"""Latency of one small VADWrapper.process_audio_data call (the per-chunk use of the reference's API)."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "cutter-vad_b200"))
import bench  # noqa: E402
from real_time_vad import VADConfig, VADWrapper  # noqa: E402

audio = bench.synth_audio(1, 16000 * 20, seed=3)[0]
w = VADWrapper(VADConfig())
w.set_callbacks(voice_start_callback=lambda: None, voice_end_callback=lambda b: None)
w.process_audio_data(audio[:16000])
for chunk in (512, 1024, 4096, 16000):
    ts = []
    for i in range(0, 16000 * 10, chunk):
        t0 = time.perf_counter()
        w.process_audio_data(audio[i:i + chunk])
        ts.append(time.perf_counter() - t0)
    frames = (chunk - 512) // 256 + 1
    print(f"chunk {chunk:6d} samples ({frames:3d} frames): call p50 {1e3 * np.median(ts):.3f} ms, p99 {1e3 * np.percentile(ts, 99):.3f} ms")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(0, 16000 * 5, 1024):
    w.process_audio_data(audio[i:i + 1024])
pr.disable(); pstats.Stats(pr).sort_stats("cumtime").print_stats(18)
