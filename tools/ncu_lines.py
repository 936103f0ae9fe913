#!/usr/bin/env python
"""Per-source-line stall samples of one kernel: joins `ncu --page source --csv` (SASS rows with sample counts, in
program order) with `nvdisasm -g` of the built library (SASS rows with //## File/line markers, same order).

    python tools/ncu_lines.py <report.ncu-rep> <kernel substring> <mangled function name> [top N]
"""
import csv
import re
import subprocess
import sys
import tempfile
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def main():
    rep, kname, mangled = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kname], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    H = rows[hdr]
    si, src = H.index("# Samples"), H.index("Source")
    sass, seen = [], set()
    for r in rows[hdr + 1:]:
        if len(r) > si and r[0].startswith("0x") and r[0] not in seen:      # a launch may be listed more than once
            seen.add(r[0])
            sass.append((r[src].strip(), float(r[si] or 0)))
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", str(ROOT / "cutter-vad_b200" / "libcvad_b200.so")], cwd=td, capture_output=True)
        cub = next(Path(td).glob("*.cubin"))
        dis = subprocess.run(["nvdisasm", "-g", str(cub)], capture_output=True, text=True).stdout.splitlines()
    lines, cur, on = [], ("?", 0), False
    for ln in dis:
        if ln.startswith(f".text.{mangled}:"):
            on = True
            continue
        if on and ln.startswith("//--------------------- "):
            break
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (Path(m.group(1)).name, int(m.group(2)))
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?) ;", ln)
        if m:
            lines.append(cur)
    if len(lines) != len(sass):
        print(f"warning: {len(lines)} disassembled instructions vs {len(sass)} profiled (library rebuilt since the capture?)")
    agg = defaultdict(float)
    for (f, l), (_, n) in zip(lines, sass):
        agg[(f, l)] += n
    tot = sum(agg.values()) or 1.0
    cache = {}
    for (f, l), n in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
        p = next((ROOT / "cutter-vad_b200" / "csrc").glob(f), None)
        if p and p not in cache:
            cache[p] = p.read_text().splitlines()
        text = cache[p][l - 1].strip()[:110] if p and l - 1 < len(cache[p]) else ""
        print(f"{100 * n / tot:5.1f}%  {f}:{l:<5d} {text}")


if __name__ == "__main__":
    main()
