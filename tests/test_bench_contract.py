"""bench.py on a box without a GPU: the reference arm (the oracle's C port on the host cores -- the one place besides
tests/ and smoke() that may execute oracle/) prints the contract's JSON line, and our own arm refuses to run instead of
falling back to anything on the CPU."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def _run(*args, timeout=300):
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, cwd=str(ROOT))


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"] == "audio_seconds_per_second" and d["unit"] == "audio-s/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] >= 3           # W >= 3 is enforced
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert "4096 concurrent 16 kHz streams" in d["config"]["workload"] and "model" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "oracle/silero_ref.c" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # 4096 streams x 32 ms per step / step time
    assert d["value"] == pytest.approx(4096 * 0.032 / (d["ms_per_step"] * 1e-3), rel=1e-6)


def test_own_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the arm runs (covered by the gpu tier and the driver's bench)")
    r = _run("--steps", "2", "--warmup", "1")
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout) and "no CPU fallback" in (r.stderr + r.stdout)
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_reference_arm_under_torchrun_prints_one_line_from_rank_0():
    """The driver launches both arms the same way for N > 1: rank 0 alone runs the CPU arm, the others exit 0."""
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", str(ROOT / "bench.py"),
                        "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["cpu_baseline"]["kind"] == "port"
