"""What the built library contains, read from its SASS (no GPU needed): sm_100a code only, the hot kernels issue their
GEMM stages on the 5th-generation tensor cores (tcgen05.mma = UTCHMMA) with accumulators read back from TMEM
(tcgen05.ld = LDTM), stream their weights by bulk async copy (cp.async.bulk = UBLKCP), and the fused one-frame kernel
carries both sides of programmatic dependent launch (griddepcontrol.launch_dependents = PREEXIT, .wait = ACQBULK).
Mnemonics: /opt/skills/guides/B200_PROFILING.md, "SASS mnemonics"."""
import re
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "cutter-vad_b200" / "libcvad_b200.so"
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.fixture(scope="module")
def sass():
    if not Path(CUOBJDUMP).exists():
        pytest.skip("cuobjdump is not installed")
    if not LIB.exists():
        import __graft_entry__ as g
        g.build()
    out = subprocess.run([CUOBJDUMP, "-sass", str(LIB)], capture_output=True, text=True, timeout=300).stdout
    funcs, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name is not None:
            funcs[name].append(line)
    return out, {k: "\n".join(v) for k, v in funcs.items()}


def _one(funcs, *needles):
    hits = [k for k in funcs if all(n in k for n in needles)]
    assert len(hits) == 1, (needles, hits)
    return funcs[hits[0]]


def test_library_holds_sm_100a_code_only(sass):
    out, _ = sass
    archs = set(re.findall(r"arch = (sm_\w+)", out))
    assert archs == {"sm_100a"}, archs


def test_fused_kernel_runs_on_tcgen05_with_bulk_copies_and_dependent_launch(sass):
    _, funcs = sass
    k = _one(funcs, "v5tc_frontend_kernel", "ILb0ELb1ELb1ELb0E")     # <DBG = false, FUSED = true, H16 = true, PROF = false>: the default
    assert k.count("UTCHMMA") >= 32                              # every GEMM stage
    assert "LDTM" in k and "UBLKCP" in k and "UTCBAR" in k       # TMEM read-back, weight stream, tcgen05.commit
    assert "PREEXIT" in k and k.count("ACQBULK") >= 3            # launch_dependents; one wait per warp role
    assert "HMMA." not in k.replace("UTCHMMA", "")               # no mma.sync path


@pytest.mark.parametrize("needles", [("v5tc_recurrent_kernel", "ILb1E"), ("v5tc_recurrent_kernel", "ILb0E"),
                                     ("resample_tc_kernel", "ILb1E"), ("v4tc_stft_kernel", "ILb0E"), ("v4tc_stft_kernel", "ILb1E"),
                                     ("v5tc_frontend_kernel", "ILb0ELb0ELb1E"), ("v5tc_frontend_kernel", "ILb0ELb1ELb0E")])
def test_every_tensor_core_kernel_issues_tcgen05(sass, needles):
    _, funcs = sass
    k = _one(funcs, *needles)
    assert "UTCHMMA" in k and "LDTM" in k and "UBLKCP" in k


def test_fft_kernels_run_on_the_fp64_pipe(sass):
    """The exact resampler and v4's STFT are double-precision FFTs (csrc/cvad_fftk.cuh): DFMA / DADD / DMUL in the SASS,
    no tensor-core instruction, and the resampler carries the primary side of programmatic dependent launch."""
    _, funcs = sass
    for needles in (("resample_fft_kernel", "ILi1E"), ("resample_fft_kernel", "ILi3E"), ("resample_fft_kernel", "ILi6E"),
                    ("v4_stft_fft_kernel",)):
        k = _one(funcs, *needles)
        assert k.count("DFMA") + k.count("DADD") + k.count("DMUL") >= 300, needles
        assert "UTCHMMA" not in k
        if needles[0] == "resample_fft_kernel":
            assert "PREEXIT" in k
