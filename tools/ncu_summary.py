#!/usr/bin/env python
"""Turn gpurun_out/*.ncu-rep + launches CSV into a committed summary under profiles/.

    python tools/ncu_summary.py <tag> <launches.csv> <prof.ncu-rep> [note] [command]
"""
import collections
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "sm__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    kn, mv = H.index("Kernel Name"), H.index("Metric Value")
    agg = collections.defaultdict(list)
    for r in rows[hdr + 1:]:
        if len(r) > mv:
            agg[r[kn]].append(float(r[mv].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    return {k: {"launches": len(v), "avg_us": round(sum(v) / len(v) / 1e3, 2), "share_of_step": round(sum(v) / tot, 4)}
            for k, v in agg.items()}


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[H.index("Kernel Name")]}
        for k in KEYS:
            if k in H:
                d[k] = f"{r[H.index(k)]} {units[H.index(k)]}".strip()
        stalls, tot = {}, 0.0
        for i, h in enumerate(H):
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                stalls[h[len("smsp__pcsamp_warps_issue_stalled_"):]] = v
                tot += v
        d["stall_share"] = {k: round(v / tot, 3) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]} if tot else {}
        res.append(d)
    return res


def main():
    tag, lcsv, rep = sys.argv[1:4]
    note = sys.argv[4] if len(sys.argv) > 4 else ""
    out = {"tag": tag, "note": note,
           "command": sys.argv[5] if len(sys.argv) > 5 else
           "python bench.py --steps 20 --warmup 3 --skip-cpu  (4096 streams x 1 frame per step)",
           "launch_list": launches(lcsv), "full_capture": raw(rep)}
    dst = ROOT / "profiles" / f"{tag}.json"
    dst.write_text(json.dumps(out, indent=1))
    (ROOT / "profiles" / f"{tag}_launches.csv").write_text(open(lcsv).read())
    print(dst)
    for k in out["full_capture"]:
        print(k["kernel"], k.get("gpu__time_duration.sum"), k.get("dram__bytes_read.sum"), k.get("dram__bytes_write.sum"),
              k.get("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"), k["stall_share"])


if __name__ == "__main__":
    main()
