"""-m gpu: the reference's own driver, run-fft.c, compiled UNMODIFIED against include/offt.h and linked to the
B200 library (offt_b200/dropin/Makefile), run the way job-test.sh:13 runs it.  Its `-v` spot print - the only
"is the answer right" check the reference has (run-fft.c:452-503) - must show the closed-form DFT of the init()
ramp, i.e. the same four numbers the unmodified reference printed over the shims (tests/test_oracle.py)."""
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
EXE = ROOT / "offt_b200" / "dropin" / "run-fft"
pytestmark = pytest.mark.gpu


def _printed(out):
    vals = []
    for line in out.splitlines():
        m = re.match(r"p 0: 0 0 (\d): (\S+) (\S+)", line)
        if m:
            vals.append(complex(float(m.group(2)), float(m.group(3))))
    return vals


def _want(N):
    return [N ** 3 * (N - 1) / 2 * 111 + 0j] + [N ** 3 / (np.exp(-2j * np.pi * k / N) - 1) for k in (1, 2, 3)]


def _need_exe():
    if not EXE.exists():
        if not Path("/root/reference/run-fft.c").exists():
            pytest.skip(f"{EXE} is built from the reference's run-fft.c, which does not exist on this box, and no prebuilt binary travelled")
        pytest.fail(f"{EXE} is not built (python -c 'import __graft_entry__ as g; g.build()')")


@pytest.mark.parametrize("flags", [[], ["-S", "1"]])
def test_run_fft_single_rank(flags):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("needs a CUDA device")
    _need_exe()
    N = 64
    res = subprocess.run([str(EXE), "-N", str(N), "-n", str(N), "-L", str(N), "-r", "2", "-m", "1", "-v", "-a", "0"] + flags,
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    got = _printed(res.stdout)
    assert len(got) == 4, res.stdout[-2000:]
    np.testing.assert_allclose(got, _want(N), rtol=1e-9, atol=1e-3)   # the driver prints 5 decimals
    assert "t_min" in res.stdout


def test_run_fft_ranks_over_offtrun():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    _need_exe()
    p = 4 if n >= 4 else 2
    N = 128
    cmd = [str(ROOT / "offt_b200" / "bin" / "offtrun"), "-n", str(p), str(EXE), "-N", str(N), "-n", str(N), "-L", str(N),
           "-r", "3", "-m", "1", "-v", "-a", "0", "-o", "-d", str(p)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    got = _printed(res.stdout)
    assert len(got) == 4, res.stdout[-2000:]
    np.testing.assert_allclose(got, _want(N), rtol=1e-9, atol=1e-2)


REF_EXE = ROOT / "oracle" / "_ref" / "run-fft"   # the unmodified reference (library included) over the MPI/FFT shims, on host cores


def _lines(out, prefixes):
    return [l.rstrip() for l in out.splitlines() if l.startswith(prefixes)]


@pytest.mark.parametrize("p,flags", [(1, []), (1, ["-S", "1", "-T", "8"]), (4, ["-o", "-d", "4"]), (4, ["-d", "2", "-t", "4", "-w", "1"]),
                                     (1, ["-R"]), (1, ["-R", "-S", "1"]),                       # real-to-complex (run-fft.c:53-55, 183)
                                     (1, ["-N", "12", "-n", "10", "-L", "9"]),                  # lengths with odd factors
                                     (1, ["-N", "24", "-n", "20", "-L", "18", "-R"]),
                                     (2, ["-N", "15", "-n", "10", "-L", "9", "-d", "2", "-o", "-V", "3"])])   # uneven split, exact counts
def test_same_stdout_as_the_reference_driver(p, flags):
    """same driver source, two libraries: the reference's own (CPU, shims) and this one (GPU).  Everything the driver
    prints except the timings must agree: echoed flags, default and final parameter lines, the M/m line, the -v values."""
    import torch
    if torch.cuda.device_count() < p:
        pytest.skip(f"needs {p} GPUs")
    _need_exe()
    if not REF_EXE.exists():
        pytest.skip("oracle/_ref/run-fft not built")
    N = 64
    args = ["-N", str(N), "-n", str(N), "-L", str(N), "-r", "1", "-m", "1", "-v", "-a", "0", "-c"] + flags   # -c: is_notest; later -N/-n/-L win
    import os
    ref = subprocess.run([str(REF_EXE)] + args, capture_output=True, text=True, timeout=300, env=dict(os.environ, OFFT_SHIM_NP=str(p)))
    cmd = [str(EXE)] + args if p == 1 else [str(ROOT / "offt_b200" / "bin" / "offtrun"), "-n", str(p), str(EXE)] + args
    got = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert ref.returncode == 0 and got.returncode == 0, got.stdout[-1500:] + got.stderr[-1500:]
    same = ("Nx ", "@ INPUT", "P1 ", "M1 ", "@ FINAL", "set p1", "allocate memory", "p1 ")
    assert _lines(got.stdout, same) == _lines(ref.stdout, same)
    a, b = _printed(got.stdout), _printed(ref.stdout)
    assert len(a) == len(b) == 4
    np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-3)


@pytest.mark.parametrize("p,flags", [(1, ["-l", "4", "-s", "3"]), (2, ["-l", "6", "-s", "0", "-o"]), (2, ["-l", "5", "-s", "2", "-O", "2", "-o"])])
def test_run_fft_with_tuning(p, flags):
    """-l N: offt_3d_init tunes before it returns (offt-compute.c:3437-3448): the driver allocates for the largest
    decomposition (run-fft.c:269-292), reads the layout afterwards, and its -v values must still be the ramp's DFT"""
    import torch
    if torch.cuda.device_count() < p:
        pytest.skip(f"needs {p} GPUs")
    _need_exe()
    N = 64
    args = ["-N", str(N), "-n", str(N), "-L", str(N), "-r", "2", "-m", "1", "-v", "-a", "0", "-c"] + flags
    cmd = [str(EXE)] + args if p == 1 else [str(ROOT / "offt_b200" / "bin" / "offtrun"), "-n", str(p), str(EXE)] + args
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "@ BEST" in res.stdout
    if (ROOT / "offt_b200" / "ah" / "_root" / "lib" / "libofft_ah.so").exists() and "-s" in flags and flags[flags.index("-s") + 1] in ("0", "2"):
        assert "Starting Harmony..." in res.stdout    # the reference's own server proposed the points (offt-tuning.c:838)
    got = _printed(res.stdout)
    assert len(got) == 4, res.stdout[-2000:]
    np.testing.assert_allclose(got, _want(N), rtol=1e-9, atol=1e-3)
