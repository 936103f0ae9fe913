"""Chained one-frame steps on the global timer: where CTAs 0, 64 and the last one of the last steps were between the
kernels (cvad_read_profile_chain).  Answers: how long after the previous step's last CTA exit does a tile start?"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "cutter-vad_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
from real_time_vad.engine import capi  # noqa: E402
from real_time_vad.engine.stream_engine import StreamEngine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
torch.cuda.set_device(0)
wl = bench.Workload(torch, 0, 0, "v5", n, 1, 16000, False, want_host=False)
eng = StreamEngine(max_streams=n, math="tc16")
L = capi.lib()
L.cvad_read_profile_chain.argtypes = [C.c_void_p, C.c_void_p]
stream = torch.cuda.Stream()
eng.set_stream(stream.cuda_stream)
eng.reset()
for i in range(50):
    eng.step_device(wl.dargs[i % wl.pool_n])
eng.sync()
eng._check(L.cvad_set_profile(eng._h, 1))
for i in range(64):
    eng.step_device(wl.dargs[i % wl.pool_n])
eng.sync()
out = np.zeros(512, np.int64)
eng._check(L.cvad_read_profile_chain(eng._h, out.ctypes.data))
m = out.reshape(8, 8, 8)[:, :3, :8]                      # [step % 8][cta 0 / 64 / last][mark]
order = np.argsort(m[:, 0, 0])                           # steps in time order
m = m[order]
t0 = m[0, 0, 0]
names = ["entry", "prologue", "dep resolved", "tile start", "tile end", "exit", "tile_live done", "at first barrier"]
print("ns since the first listed step's CTA-0 entry; rows = consecutive steps")
for b, nm in enumerate(("CTA 0 (scheduled early)", "CTA 64", "last CTA")):
    print(nm)
    for k in range(8):
        print("   ", "  ".join(f"{names[j]} {int(m[k, b, j] - t0):8d}" for j in (0, 1, 2, 6, 7, 3, 4, 5)))
per = np.diff(m[:, 2, 4]).mean()
print(f"step period (tile end to tile end, last CTA): {per:.0f} ns; tile {np.mean(m[:, 2, 4] - m[:, 2, 3]):.0f} ns; "
      f"previous step's tile end -> this step's tile start (last CTA): {np.mean(m[1:, 2, 3] - m[:-1, 2, 4]):.0f} ns; "
      f"of which exit -> next entry {np.mean(m[1:, 2, 0] - m[:-1, 2, 5]):.0f}, entry -> prologue {np.mean(m[:, 2, 1] - m[:, 2, 0]):.0f}, "
      f"prologue -> dep resolved {np.mean(m[:, 2, 2] - m[:, 2, 1]):.0f}, dep resolved -> tile start {np.mean(m[:, 2, 3] - m[:, 2, 2]):.0f}")

# the tile's phase marks (clock64, CTA 0) of the LAST chained step: the same table tools/tc_phase_profile.py prints for a
# host-buffer step, here with the audio prefetched into L2 by the early CTAs and the state left in L2 by the previous step
prof = np.zeros(128, np.int64)
eng._check(L.cvad_read_profile(eng._h, prof.ctypes.data))
fe = prof[0:12] - prof[0]
names = ["tile start", "loader done", "stft acc", "stft epi", "enc0 acc", "enc0 epi", "enc1 acc", "enc1 epi", "enc2 acc",
         "enc2 epi", "enc3 acc", "enc3 epi"]
print("chained step, CTA 0, cycles since tile start (delta):")
for i, nm in enumerate(names):
    print(f"  {nm:12s} {fe[i]:8d} {fe[i] - (fe[i - 1] if i else 0):8d}")
print("  state loaded", prof[12] - prof[0], "| loader gated", prof[20] - prof[0], "amax pushed", prof[15] - prof[0], "| x ready", prof[43] - prof[0],
      "first gate ready", prof[13] - prof[0], "i,f,g done", prof[21] - prof[0], "cell computed", prof[22] - prof[0],
      "state stored + barrier", prof[23] - prof[0], "cell done", prof[14] - prof[0])
