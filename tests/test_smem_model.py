"""CPU: shared-memory bank behaviour of the configured kernels, from tools/smem_model.py (which mirrors the index
algebra of fft_kernels.cuh): contiguous-row launches must be conflict-free for every length, strided launches within
the recorded bounds.  Guards the PAD constants of csrc/fft_configs.h against accidental edits."""
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))
from smem_model import Cfg, score  # noqa: E402

# worst strided case the current pads leave (wavefronts per request, 1.0 = conflict-free); DESIGN.md section 9 lists
# the 1024-point complex128 figure (1.25 at 4 columns) as a known cost of pads shared between the two lane orders
STRIDED_BOUND = {"c128": 1.25, "c64": 1.75}


def configs():
    txt = (ROOT / "offt_b200" / "csrc" / "fft_configs.h").read_text()
    defs = dict(re.findall(r"#define (OFFTB_CFG_\d+) ([\d, ]+)", txt))
    rows = []
    for m in re.finditer(r"X\((\d+), ([\d, ]+?), (?:\d+|OFFTB_MAXT_\d+), (?:\d+|OFFTB_MINB_\d+)\)", txt):
        rows.append((int(m.group(1)), [int(x) for x in m.group(2).split(",")]))
    for m in re.finditer(r"OFFTB_APPLY\(X, (\d+), (OFFTB_CFG_\d+)", txt):
        rows.append((int(m.group(1)), [int(x) for x in defs[m.group(2)].split(",")]))
    return sorted(rows)


def test_configured_pads():
    rows = configs()
    assert [n for n, _ in rows] == [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192]
    for N, vals in rows:
        E, R = vals[0], [r for r in vals[1:5] if r > 1] or [vals[1]]
        prod = 1
        for r in R:
            prod *= r
        assert prod == N and all(E % r == 0 for r in R)
        if len(R) < 2 or N > 2048:      # single-stage lengths exchange nothing; the two longest take minutes to model
            continue
        for G, name, pads in ((8, "c128", vals[5:8]), (16, "c64", vals[8:11])):
            cfg = Cfg(N, E, R, pads)
            assert score(cfg, 1, False, G) == 1.0, f"N={N} {name}: contiguous-row launches must be conflict-free"
            for C in (4, 8):
                assert score(cfg, C, True, G) <= STRIDED_BOUND[name], f"N={N} {name} strided, {C} columns"
