"""CPU: the rank-local layout descriptor of the library (offtb_comm_fill = what offt_comm_malloc hands the caller,
offt-compute.c:57-315) against (a) the descriptors the UNMODIFIED reference produced for every rank of every golden
fixture - even and uneven splits, slabs and pencils, all three output layouts - and (b) the oracle over a sweep."""
import itertools
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from conftest import golden_names, load_golden  # noqa: E402

KEYS = "istart isize istride ostart osize ostride".split()


@pytest.mark.parametrize("name", golden_names())
def test_descriptor_matches_reference_fixture(name):
    import offt_b200 as ob
    g = load_golden(name)
    Nx, Ny, Nz = g["N"]
    S = g["params"][23]
    for b in g["boxes"]:
        got = ob.comm_box(Nx, Ny, Nz, g["p"], b.p1, b.rank, S=S, is_equalxy=g["is_equalxy"])
        assert (got["p1"], got["p2"]) == (b.p1, b.p2)
        for k in KEYS:
            assert tuple(got[k]) == tuple(getattr(b, k)), f"{name} rank {b.rank} {k}"
        assert ob.alloc_elems(Nx, Ny, Nz, g["p"], b.p1) == b.alloc


def test_descriptor_matches_oracle_sweep(oracle):
    import offt_b200 as ob
    fields = "p1 p2 M1 M2 M3 M4 F1 F2 F3 F4 m1 m2 m3 m4 b1 b2 b3 b4".split()
    n = 0
    for (Nx, Ny, Nz), p in itertools.product([(16, 8, 32), (12, 10, 9), (64, 64, 64), (20, 36, 28), (1024, 1024, 1024)], [1, 2, 3, 4, 6, 8]):
        for p1 in [d for d in range(1, p + 1) if p % d == 0]:
            p2 = p // p1
            if p1 > min(Nx, Ny) or p2 > min(Ny, Nz):
                continue
            for S, eq in ((0, 0), (1, 0), (0, 1)):
                for rank in range(p):
                    a = ob.comm_box(Nx, Ny, Nz, p, p1, rank, S=S, is_equalxy=eq)
                    b = oracle.comm(Nx, Ny, Nz, p, p1, rank, S, eq)
                    for k in fields + KEYS:
                        assert tuple(a[k]) == tuple(b[k]) if k in KEYS else a[k] == b[k], (Nx, Ny, Nz, p, p1, rank, S, eq, k)
                    n += 1
    assert n > 500
