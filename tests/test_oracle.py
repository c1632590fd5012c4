"""CPU suite: pins the oracle (oracle/offt_oracle.c) before anything trusts it.

1. against the golden fixtures generated from the UNMODIFIED reference
   (tests/golden/make_golden.py) - bit-exact, raw in-place arrays included;
2. against the closed form of run-fft.c's input ramp (run-fft.c:46-61), i.e. the
   four numbers the reference's only self-check (-v, run-fft.c:452-503) prints;
3. against numpy.fft.fftn as an independent arithmetic check (<= 1e-14 rel-L2);
4. live against oracle/_ref/ref_dump when it is present (build container).
"""
import numpy as np
import pytest

from conftest import golden_names, load_golden
from oracle import oracle as O


@pytest.mark.parametrize("name", golden_names())
def test_restatement_matches_reference_fixture(oracle, name):
    g = load_golden(name)
    N, p = g["N"], g["p"]
    v = oracle.resolve_params(*N, p, g["custom"])
    assert v == g["params"], "params_set_default + custom override differ from the reference"
    grid = O.grid_values(g["seed"], *N)
    mine = oracle.execute(grid, p, v, g["is_oned"], g["is_equalxy"])
    for a, b in zip(mine, g["boxes"]):
        assert (a.p1, a.p2, a.istart, a.isize, a.istride, a.ostart, a.osize, a.ostride, a.alloc) == \
               (b.p1, b.p2, b.istart, b.isize, b.istride, b.ostart, b.osize, b.ostride, b.alloc)
        assert np.array_equal(a.data, b.data), f"rank {a.rank}: in-place array differs from the reference"
    spec = O.gather_output(mine)
    assert not np.isnan(spec).any()
    assert O.rel_l2(spec, np.fft.fftn(grid)) < 1e-14


@pytest.mark.parametrize("N,p,p1,oned,S", [((64, 64, 64), 4, 4, 1, 0), ((16, 32, 8), 4, 2, 0, 1), ((8, 8, 8), 1, 1, 0, 0)])
def test_ramp_known_answer(oracle, N, p, p1, oned, S):
    """closed form of the DFT of in[x,y,z] = z + 10y + 100x."""
    Nx, Ny, Nz = N
    v = oracle.resolve_params(*N, p, {O.P1: p1, O.S: S})
    spec = O.gather_output(oracle.execute(O.ramp_values(*N), p, v, oned, 0))
    tot = Nx * Ny * Nz
    assert spec[0, 0, 0] == pytest.approx(tot * ((Nz - 1) / 2 + 10 * (Ny - 1) / 2 + 100 * (Nx - 1) / 2), rel=1e-13)
    k = np.arange(1, Nz)
    np.testing.assert_allclose(spec[0, 0, 1:], tot / (np.exp(-2j * np.pi * k / Nz) - 1), rtol=1e-12)
    k = np.arange(1, Ny)
    np.testing.assert_allclose(spec[0, 1:, 0], 10 * tot / (np.exp(-2j * np.pi * k / Ny) - 1), rtol=1e-12)
    k = np.arange(1, Nx)
    np.testing.assert_allclose(spec[1:, 0, 0], 100 * tot / (np.exp(-2j * np.pi * k / Nx) - 1), rtol=1e-12)
    off_axis = spec.copy()
    off_axis[0, 0, :] = 0; off_axis[0, :, 0] = 0; off_axis[:, 0, 0] = 0
    assert np.abs(off_axis).max() < 1e-9 * abs(spec[0, 0, 0])


def test_reference_verbose_print_values():
    """the four numbers `run-fft -N 64 -n 64 -L 64 -v -a 0 -o -d 4` printed when the unmodified
    reference ran over the shims in the build container (recorded 2026-10-18)."""
    printed = [(916586496.00000, 0.00000), (-131072.00000, 2668031.85254),
               (-131072.00000, 1330796.34904), (-131072.00000, 883615.64968)]
    N = 64
    want = [N ** 3 * (N - 1) / 2 * 111 + 0j] + [N ** 3 / (np.exp(-2j * np.pi * k / N) - 1) for k in (1, 2, 3)]
    for (re, im), w in zip(printed, want):
        assert abs(complex(re, im) - w) < 1e-4


def test_delta_and_plane_wave(oracle):
    N, p = (8, 16, 32), 4
    v = oracle.resolve_params(*N, p, {O.P1: 2})
    g = np.zeros(N, dtype=np.complex128); g[0, 0, 0] = 1
    assert np.allclose(O.gather_output(oracle.execute(g, p, v)), 1.0)
    x, y, z = np.meshgrid(np.arange(N[0]), np.arange(N[1]), np.arange(N[2]), indexing="ij")
    a, b, c = 3, 5, 7
    g = np.exp(2j * np.pi * (a * x / N[0] + b * y / N[1] + c * z / N[2]))
    spec = O.gather_output(oracle.execute(g, p, v))
    want = np.zeros(N, dtype=np.complex128); want[a, b, c] = np.prod(N)
    assert np.abs(spec - want).max() < 1e-9


def test_dft_rows_matches_numpy(oracle):
    rng = np.random.default_rng(0)
    for n in (2, 8, 30, 64, 97, 256):
        a = (rng.standard_normal((5, n)) + 1j * rng.standard_normal((5, n)))
        b = a.copy()
        oracle.dft_rows(b, n, 1, n, 5, -1)
        assert O.rel_l2(b, np.fft.fft(a, axis=1)) < 1e-14
        b = np.ascontiguousarray(a.T.copy())         # rows along the slow axis: stride 5
        oracle.dft_rows(b, n, 5, 1, 5, +1)
        assert O.rel_l2(b.T, np.fft.ifft(a, axis=1) * n) < 1e-14


@pytest.mark.skipif(not O.have_reference(), reason="oracle/_ref/ref_dump not built (needs /root/reference)")
@pytest.mark.parametrize("N,p,oned,eq,custom", [
    ((16, 16, 16), 4, 0, 0, {O.P1: 4}),
    ((16, 16, 16), 2, 1, 1, {O.P1: 2, O.T2: 4, O.W2: 0}),
    ((24, 20, 12), 4, 0, 0, {O.P1: 2, O.V: 3, O.T1: 5}),
    ((8, 16, 64), 8, 0, 0, {O.P1: 2, O.S: 1, O.RY: 7}),
])
def test_restatement_matches_live_reference(oracle, N, p, oned, eq, custom):
    boxes, _ = O.run_reference(*N, p, seed=11, is_oned=oned, is_equalxy=eq, params=custom)
    v = oracle.resolve_params(*N, p, custom)
    assert v == boxes[0].params
    mine = oracle.execute(O.grid_values(11, *N), p, v, oned, eq)
    for a, b in zip(mine, boxes):
        assert np.array_equal(a.data, b.data)


def test_params_range_and_feasibility(oracle):
    """value grids of params_range_setup (offt-compute.c:2998-3093) and is_infeasible_point."""
    r = oracle.params_range(1024, 1024, 1024, 8)
    assert r[O.P1] == [1, 2, 4, 8]
    assert r[O.T1] == [2 ** i for i in range(11)]
    assert r[O.W1] == list(range(11)) and r[O.V] == [0, 1, 2, 3] and r[O.S] == [0, 1]
    assert r[O.FZ][:3] == [0, 1, 2]
    r = oracle.params_range(12, 10, 9, 4)
    assert r[O.T1] == [1, 2, 4, 8, 12] and r[O.T2] == [1, 2, 4, 8, 9]
    v = oracle.params_default(1024, 1024, 1024, 8, 0, 1)
    # defaults are derived with P1 = floor(sqrt(p)) = 2
    assert (v[O.P1], v[O.T1], v[O.W1], v[O.T2], v[O.W2], v[O.RY]) == (2, 32, 2, 16, 2, 5)
    # ... and at 1024^3 they break the reference's own BUFFER_SIZE_LIMIT rule (offt-tuning.c:170)
    assert oracle.is_infeasible(1024, 1024, 1024, 8, v) == (1, O.W1)
    v = oracle.params_default(256, 256, 256, 4, 0, 1)
    assert oracle.is_infeasible(256, 256, 256, 4, v) == (0, -1)
    bad = list(v); bad[O.T1] = 4096
    assert oracle.is_infeasible(256, 256, 256, 4, bad) == (1, O.T1)
    bad = list(v); bad[O.W2] = 10; bad[O.T2] = 128
    assert oracle.is_infeasible(256, 256, 256, 4, bad) == (1, O.W2)
