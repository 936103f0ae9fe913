"""The C ABI from plain C: examples/c_abi_demo.c includes only include/cutter_vad_b200.h, links libcvad_b200.so with gcc
and drives engine + stream feeder (no Python, no torch in the process).  CPU: it compiles, links, and refuses to run
without a B200 (no fallback).  GPU: 64 streams of a gated tone give exactly one voice segment each."""
import shutil
import subprocess
import sys

import pytest

from conftest import PKG, ROOT


def _build(tmp_path):
    exe = tmp_path / "c_abi_demo"
    subprocess.run(["gcc", "-O2", "-Wall", "-Werror", f"-I{ROOT / 'include'}", str(ROOT / "examples" / "c_abi_demo.c"),
                    f"-L{PKG}", "-lcvad_b200", f"-Wl,-rpath,{PKG}", "-lm", "-o", str(exe)], check=True)
    blob = tmp_path / "v5.f32"
    subprocess.run([sys.executable, str(ROOT / "tools" / "export_weights.py"), "v5", str(blob)], check=True, capture_output=True)
    assert blob.stat().st_size == 309633 * 4
    return exe, blob


@pytest.mark.skipif(shutil.which("nvidia-smi") is not None, reason="only meaningful without a GPU")
def test_c_program_links_and_fails_loudly_without_a_gpu(tmp_path):
    exe, blob = _build(tmp_path)
    r = subprocess.run([str(exe), str(blob)], capture_output=True, text=True)
    assert r.returncode == 3 and "no sm_100 device" in r.stderr


@pytest.mark.gpu
def test_c_program_detects_one_segment_per_stream(tmp_path):
    exe, blob = _build(tmp_path)
    r = subprocess.run([str(exe), str(blob)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "voice starts 64, ends 64" in r.stdout
    assert "stream 0: VOICE_START" in r.stdout and "stream 0: VOICE_END" in r.stdout
