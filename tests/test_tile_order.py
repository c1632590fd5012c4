"""CPU: the order in which a phase visits its tiles (plan.cu, tile_visited) against an interval model of what the two
launches of a tile read and write when the exchange runs in place (_S_ = 1: the array between the phases is the caller's).

Phase 1, forward: the writer of a tile reads x planes of the caller's layout (plane stride istride[0]); the reader writes
the same number of planes of the x-y-z_local layout (plane stride M3*M4*p1).  Backward: the other way round.  A reader
may run as soon as its own tile has arrived, i.e. while the writers of all LATER visits have not read their planes yet
(with W slots in flight the writers of the next W visits may already have, the others certainly have not) - so what the
reader of visit v writes must not touch what the writers of visits > v read.  With an uneven division the two strides
differ and only one direction of travel satisfies that; the backward transform of 27x20x45 on 8 ranks went the wrong way
until round 2 (profiles/r02_exchange_ab.md)."""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import offt_b200 as ob  # noqa: E402
from offt_b200.binding import lib  # noqa: E402


def visits(nb, phase, inverse):
    order = [lib.offtb_tile_visited(nb, v, phase, int(inverse)) for v in range(nb)]
    assert sorted(order) == list(range(nb))                       # every tile exactly once
    assert lib.offtb_tile_visited(nb, nb, phase, int(inverse)) == -1 and lib.offtb_tile_visited(nb, -1, phase, int(inverse)) == -1
    return order


def test_forward_goes_up_and_backward_phase_1_goes_down():
    for nb in (1, 2, 5, 16):
        assert visits(nb, 1, False) == list(range(nb))
        assert visits(nb, 2, False) == list(range(nb))
        assert visits(nb, 2, True) == list(range(nb))
        assert visits(nb, 1, True) == list(range(nb))[::-1]


def plane_span(stride, x0, n):
    """elements [lo, hi) that n planes starting at plane x0 can touch when planes are `stride` elements apart"""
    return (x0 * stride, (x0 + n) * stride)


CASES = [
    # N, p, p1: slab 1 x p and pencils whose y rows do not divide evenly over p2 (istride[0] > M3*M4*p1), and even ones
    ((27, 20, 45), 8, 1), ((27, 11, 45), 2, 1), ((27, 21, 45), 4, 1), ((12, 11, 15), 4, 2), ((12, 13, 15), 8, 2), ((20, 12, 18), 6, 2),
    ((30, 42, 70), 7, 1), ((64, 64, 64), 4, 1), ((64, 64, 64), 8, 2), ((96, 80, 48), 8, 4), ((1024, 1024, 1024), 8, 1),
]


@pytest.mark.parametrize("N,p,p1", CASES)
@pytest.mark.parametrize("T1", [1, 2, 4, 5, 64])
def test_in_place_phase_1_never_overwrites_what_a_later_tile_reads(N, p, p1, T1):
    skewed = 0
    for rank in range(p):
        box = ob.comm_box(*N, p, p1, rank, S=1)
        isx = box["istride"][0]                         # plane stride of the caller's layout
        dX = box["M3"] * box["M4"] * box["p1"]          # plane stride between the phases (plan.cu, dims_of)
        assert isx >= dX                                # what the rule relies on
        skewed += isx != dX
        planes = box["m1"]
        nb = (planes + T1 - 1) // T1
        for inverse in (False, True):
            order = visits(nb, 1, inverse)
            rd_stride, wr_stride = (dX, isx) if inverse else (isx, dX)
            for v, tile in enumerate(order):
                n = min(T1, planes - tile * T1)
                w_lo, w_hi = plane_span(wr_stride, tile * T1, n)
                for later in order[v + 1:]:
                    r_lo, r_hi = plane_span(rd_stride, later * T1, min(T1, planes - later * T1))
                    assert w_hi <= r_lo or r_hi <= w_lo, (rank, inverse, tile, later, (w_lo, w_hi), (r_lo, r_hi))
    if N in ((27, 20, 45), (27, 11, 45), (12, 13, 15)):   # 12x13x15 on 2x4: 16 rows of y in the caller's planes, 14 between the phases
        assert skewed                                   # these are the cases that need the rule


def test_the_opposite_order_would_collide_on_the_skewed_case():
    """the model itself can see the bug: 27x20x45 on 8 ranks, backward, ascending tiles"""
    box = ob.comm_box(27, 20, 45, 8, 1, 0, S=1)
    isx, dX, planes, T1 = box["istride"][0], box["M3"] * box["M4"] * box["p1"], box["m1"], 4
    nb = (planes + T1 - 1) // T1
    assert isx > dX and nb > 1
    clash = False
    order = list(range(nb))                             # ascending, as before the fix
    for v, tile in enumerate(order):
        w_lo, w_hi = plane_span(isx, tile * T1, min(T1, planes - tile * T1))
        for later in order[v + 1:]:
            r_lo, r_hi = plane_span(dX, later * T1, min(T1, planes - later * T1))
            clash |= not (w_hi <= r_lo or r_hi <= w_lo)
    assert clash
