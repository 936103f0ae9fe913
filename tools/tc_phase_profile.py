"""Phase timeline of CTA 0 of the tensor-core kernels (clock64 marks; cvad_set_profile)."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT / "cutter-vad_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
from conftest import synth_streams  # noqa: E402
from real_time_vad.engine.stream_engine import StreamEngine  # noqa: E402

n, T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, int(sys.argv[2]) if len(sys.argv) > 2 else 1
eng = StreamEngine(max_streams=n, math=sys.argv[3] if len(sys.argv) > 3 else "tc")
x = np.tile(synth_streams(64, 512 * T, seed=1), (n // 64, 1))
for _ in range(3):
    eng.step(x)
eng._check(eng._L.cvad_set_profile(eng._h, 1))
eng.step(x)
out = np.zeros(128, np.int64)
eng._check(eng._L.cvad_read_profile(eng._h, out.ctypes.data))
fe = out[0:12] - out[0]
mm = out[32:42] - out[0]
names = ["tile start", "loader done", "stft acc", "stft epi", "enc0 acc", "enc0 epi", "enc1 acc", "enc1 epi", "enc2 acc",
         "enc2 epi", "enc3 acc", "enc3 epi"]
print("front end, epilogue thread 0 (cycles since tile start, delta):")
for i, nm in enumerate(names):
    print(f"  {nm:12s} {fe[i]:8d} {fe[i] - (fe[i - 1] if i else 0):8d}")
print("front end, MMA warp: (act_ready seen, layer issued) x 5:", mm.tolist())
if out[12]:
    print("fused: state loaded", out[12] - out[0], "| hh issued", out[42] - out[0], "| x ready", out[43] - out[0], "ih issued", out[44] - out[0], "| gates ready", out[13] - out[0], "cell done", out[14] - out[0])
if out[15]:
    print("tc16 marks: loader gated", out[20] - out[0], "amax pushed", out[15] - out[0], "barrier", out[16] - out[0],
          "| stft computed", out[17] - out[0], "amax pushed", out[18] - out[0], "barrier", out[19] - out[0])
if out[21]:
    print("fused cell marks: gates in registers", out[21] - out[0], "cell computed", out[22] - out[0], "state transposed + barrier", out[23] - out[0],
          "state stored", out[24] - out[0], "cell done", out[14] - out[0])
r0 = out[64]
print("recurrent epilogue: start 0, state loaded", out[65] - r0)
for j in range(min(T, 8)):
    print(f"  frame {j}: acc_ready {out[66 + 3 * j] - r0}, h written {out[67 + 3 * j] - r0}, frame done {out[68 + 3 * j] - r0}"
          f" | mma: x ready {out[96 + 3 * j] - r0}, h ready {out[97 + 3 * j] - r0}, issued {out[98 + 3 * j] - r0}")
g = out[120:126] - out[120]
print("globaltimer ns: FE entry 0, FE prologue done", g[1], "FE exit", g[2], "| REC entry", g[3], "REC prologue done", g[4], "REC exit", g[5])
if out[126] and out[127] and out[14]:
    # the same two points of the tile on both clocks: SM clock actually run, and where the tile sits in the CTA's life
    ns = int(out[127] - out[126])
    cyc = int(out[14] - out[0])
    print(f"tile on the global timer: start at {int(out[126] - out[120])} ns after CTA entry, {ns} ns long = {cyc} cycles "
          f"-> {cyc / max(ns, 1) * 1e3:.0f} MHz; CTA exits {int(out[122] - out[127])} ns after the tile's last mark")
