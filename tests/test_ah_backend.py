"""CPU: the optional Active Harmony back end of the tuner (offt_b200/ah): the reference's own hserver + session-core +
patched nm.so, compiled unmodified, driven through offt_b200/ah/ah_glue.c exactly as the reference's ah_tuning drives
them (offt-tuning.c:773-854): a 24-variable session over grid indices, the user simplex file, fetch / report, best.
The objective here is a synthetic bowl - no FFT, no GPU - so this checks the plumbing (server launch on localhost,
session, strategy plug-in, protocol) wherever the back end was built."""
import ctypes as C
import os
import random
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
AH_ROOT = ROOT / "offt_b200" / "ah" / "_root"


@pytest.mark.skipif(not (AH_ROOT / "lib" / "libofft_ah.so").exists(), reason="offt_b200/ah/_root not built (needs the reference tree at build time)")
@pytest.mark.parametrize("strategy", [0, 2])   # nm.so from the user simplex; random.so
def test_harmony_session_minimises_a_bowl(strategy, tmp_path):
    L = C.CDLL(str(AH_ROOT / "lib" / "libofft_ah.so"))
    L.offtb_ah_error.restype = C.c_char_p
    L.offtb_ah_report.argtypes = [C.c_double]
    n = 24
    sizes = (C.c_int * n)(*([8] * n))
    rng = random.Random(3)
    uv = tmp_path / "uv"
    uv.write_text("".join(" ".join(str(rng.randrange(8)) for _ in range(n)) + " \n" for _ in range(n + 1)))
    port = 21000 + os.getpid() % 20000 + strategy
    rc = L.offtb_ah_open(str(AH_ROOT).encode(), n, sizes, strategy, str(uv).encode() if strategy == 0 else None, port)
    assert rc == 0, L.offtb_ah_error()
    try:
        idx = (C.c_long * n)()
        first = best = None
        for _ in range(300):
            if L.offtb_ah_converged() == 1:
                break
            assert L.offtb_ah_fetch(idx) >= 0
            v = list(idx)
            assert all(0 <= a < 8 for a in v)
            f = float(sum((a - 3) ** 2 for a in v))
            first = f if first is None else first
            best = f if best is None else min(best, f)
            assert L.offtb_ah_report(f) == 0
        assert best < first
        if strategy == 0:
            assert best <= 0.2 * first          # Nelder-Mead actually descends
        assert L.offtb_ah_best(idx) >= 0
    finally:
        L.offtb_ah_close()
