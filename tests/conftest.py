import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    return O.Oracle()


def load_golden(name, sub=""):
    """fixture written by tests/golden/make_golden.py from the unmodified reference"""
    import numpy as np
    from oracle import oracle as O
    z = np.load(ROOT / "tests" / "golden" / sub / f"{name}.npz")
    N = tuple(int(v) for v in z["N"])
    p = int(z["p"])
    boxes, off = [], 0
    for r, d in enumerate(z["desc"]):
        d = [int(v) for v in d]
        alloc = d[20]
        boxes.append(O.RankBox(p=p, rank=r, N=N, p1=d[0], p2=d[1], istart=tuple(d[2:5]), isize=tuple(d[5:8]),
                               istride=tuple(d[8:11]), ostart=tuple(d[11:14]), osize=tuple(d[14:17]),
                               ostride=tuple(d[17:20]), alloc=alloc, params=[int(v) for v in z["params"]],
                               data=z["data"][off:off + alloc]))
        off += alloc
    custom = {int(k): int(v) for k, v in z["custom"]}
    return dict(N=N, p=p, is_oned=int(z["is_oned"]), is_equalxy=int(z["is_equalxy"]), seed=int(z["seed"]),
                custom=custom, params=[int(v) for v in z["params"]], boxes=boxes)


def golden_names(sub=""):
    return sorted(f.stem for f in (ROOT / "tests" / "golden" / sub).glob("*.npz"))
