"""-m gpu: the CUDA path through the C-ABI library against the oracle.

Tolerances are BASELINE.json's: relative L2 error <= 1e-12 in double, <= 1e-5 in single
(single is compared with the double oracle).  Layouts are compared through
ostart/osize/ostride, exactly where the reference driver reads its output (run-fft.c:477-478).
"""
import numpy as np
import pytest

from conftest import golden_names, load_golden
from gpu_helpers import gather_input, gpu_forward, local_world
from oracle import oracle as O

pytestmark = pytest.mark.gpu
TOL = {64: 1e-12, 32: 1e-5}
P = O  # parameter indices


def _torch():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device; the library has no CPU path")
    return torch


# ---------------------------------------------------------------- the 1-D kernels alone
@pytest.mark.parametrize("bits", [64, 32])
@pytest.mark.parametrize("n", [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192])
def test_rows_contiguous_and_strided(oracle, n, bits):
    torch = _torch()
    import offt_b200 as ob
    rng = np.random.default_rng(n)
    rows = 16
    a = (rng.uniform(-1, 1, (rows, n)) + 1j * rng.uniform(-1, 1, (rows, n)))
    want = a.copy()
    oracle.dft_rows(want, n, 1, n, rows, -1)
    cdt = torch.complex128 if bits == 64 else torch.complex64
    with local_world(1):
        d = torch.from_numpy(a).to("cuda").to(cdt).contiguous()
        ob.fft_rows(d, n, 1, n, rows, sign=-1, bits=bits)
        assert O.rel_l2(d.cpu().numpy(), want) < TOL[bits]
        # same rows laid out along the slow axis: element j of row h at h + j*rows
        d = torch.from_numpy(np.ascontiguousarray(a.T)).to("cuda").to(cdt).contiguous()
        ob.fft_rows(d, n, rows, 1, rows, sign=-1, bits=bits)
        assert O.rel_l2(d.cpu().numpy().T, want) < TOL[bits]
        # backward transform of the result returns n * input
        ob.fft_rows(d, n, rows, 1, rows, sign=+1, bits=bits)
        assert O.rel_l2(d.cpu().numpy().T / n, a) < TOL[bits]


@pytest.mark.parametrize("n", [16, 64, 512, 1024])
def test_raw_maps_split_transpose_ry(oracle, n):
    """the address maps that fuse pack/unpack: split transform index, batch digits (also
    non power-of-two), transposing store, the Ry rule's copy path."""
    torch = _torch()
    import offt_b200 as ob
    rng = np.random.default_rng(7 * n)
    B0, B1, B2 = 8, 3, 2            # batch digits; B1 = 3 is a ragged tile
    nb = B0 * B1 * B2
    src = rng.uniform(-1, 1, (B2, B1, n, B0)) + 1j * rng.uniform(-1, 1, (B2, B1, n, B0))   # column index fastest
    want = np.fft.fft(src, axis=2)
    with local_world(1):
        d_in = torch.from_numpy(src).to("cuda").contiguous()
        # (1) strided in place
        im = [0, 0, 0, B0, B0, 1, B1, n * B0, B1 * n * B0]
        d = d_in.clone()
        ob.fft_launch_raw(d, d, n, nb, im, im, load_cfast=1, store_cfast=1)
        assert O.rel_l2(d.cpu().numpy(), want) < 1e-12
        # (2) transposing store: out[b2][b1][b0][n] (transform index contiguous)
        om = [0, 0, 0, 1, B0, n, B1, n * B0, B1 * n * B0]
        out = torch.zeros_like(d_in)
        ob.fft_launch_raw(d_in, out, n, nb, im, om, load_cfast=1, store_cfast=0)
        got = out.cpu().numpy().reshape(B2, B1, B0, n).transpose(0, 1, 3, 2)
        assert O.rel_l2(got, want) < 1e-12
        # (3) and back: load contiguous rows, store strided
        back = torch.zeros_like(d_in)
        ob.fft_launch_raw(out, back, n, nb, om, im, sign=+1, load_cfast=0, store_cfast=1)
        assert O.rel_l2(back.cpu().numpy() / n, src) < 1e-12
        # (4) split transform index on the store side: n -> (n // q, n % q) with block stride
        q = n // 4
        om = [0, q, B2 * B1 * q * B0, B0, B0, 1, B1, q * B0, B1 * q * B0]
        out = torch.zeros_like(d_in)
        ob.fft_launch_raw(d_in, out, n, nb, im, om, load_cfast=1, store_cfast=1)
        got = out.cpu().numpy().reshape(4, B2, B1, q, B0).transpose(1, 2, 0, 3, 4).reshape(B2, B1, n, B0)
        assert O.rel_l2(got, want) < 1e-12
        # (5) Ry rule on digit 2 (x): x = 7 + b2, transform only if x % 10 < 8 -> b2 = 0 only
        d = d_in.clone()
        ob.fft_launch_raw(d, d, n, nb, im, im, load_cfast=1, store_cfast=1, ry=(2, 7, 0, 8))
        got = d.cpu().numpy()
        assert O.rel_l2(got[0], want[0]) < 1e-12
        assert np.array_equal(got[1], src[1])


# ---------------------------------------------------------------- whole plans vs the oracle
CASES = [
    # N, p, is_oned, is_equalxy, custom
    ((16, 8, 32), 1, 0, 0, {P.P1: 1}),
    ((16, 8, 32), 1, 1, 0, {P.P1: 1, P.S: 1}),
    ((16, 16, 8), 1, 0, 1, {P.P1: 1}),
    ((16, 8, 32), 4, 1, 0, {P.P1: 4}),
    ((16, 8, 32), 4, 1, 0, {P.P1: 4, P.S: 1}),
    ((16, 8, 32), 4, 1, 0, {P.P1: 1}),
    ((16, 8, 32), 4, 1, 0, {P.P1: 1, P.S: 1}),
    ((16, 16, 8), 4, 1, 1, {P.P1: 1}),
    ((16, 16, 8), 4, 1, 1, {P.P1: 4}),
    ((16, 16, 16), 4, 0, 0, {P.P1: 2}),
    ((16, 16, 16), 4, 0, 0, {P.P1: 2, P.S: 1}),
    ((16, 16, 16), 4, 0, 1, {P.P1: 2}),
    ((16, 16, 16), 4, 0, 0, {P.P1: 4}),
    ((16, 16, 16), 4, 0, 0, {P.P1: 1, P.S: 1}),
    ((32, 64, 16), 8, 0, 0, {P.P1: 2, P.T1: 4, P.T2: 2}),
    ((32, 64, 16), 8, 0, 0, {P.P1: 4, P.S: 1, P.RY: 3}),
    ((32, 32, 32), 4, 0, 0, {P.P1: 2, P.T1: 3, P.T2: 5, P.W1: 1, P.W2: 3}),     # ragged last tiles
    ((32, 32, 32), 4, 0, 0, {P.P1: 2, P.T1: 16, P.T2: 16, P.W1: 0, P.W2: 0, P.RY: 0}),
    ((32, 32, 32), 4, 0, 0, {P.P1: 2, P.T1: 1, P.T2: 1, P.W1: 10, P.W2: 10, P.RY: 10, P.S: 1}),
    ((64, 128, 256), 8, 0, 0, {P.P1: 2}),
    ((256, 64, 128), 8, 1, 0, {P.P1: 8}),
    ((128, 128, 128), 2, 1, 0, {P.P1: 2, P.S: 1}),
]


@pytest.mark.parametrize("N,p,oned,eq,custom", CASES)
def test_plan_matches_oracle(oracle, N, p, oned, eq, custom):
    _torch()
    grid = O.grid_values(5, *N)
    v = oracle.resolve_params(*N, p, custom)
    want = oracle.execute(grid, p, v, oned, eq)
    got, launches, back = gpu_forward(grid, p, custom, oned, eq, inverse_too=True)
    assert launches > 0
    for a, b in zip(got, want):
        assert a.params == b.params
        assert (a.istart, a.isize, a.istride, a.ostart, a.osize, a.ostride, a.alloc) == \
               (b.istart, b.isize, b.istride, b.ostart, b.osize, b.ostride, b.alloc)
    A, B = O.gather_output(got), O.gather_output(want)
    assert not np.isnan(A).any()
    assert O.rel_l2(A, B) < 1e-12
    # forward -> backward returns N * input in the input layout
    rt = gather_input(got, back) / np.prod(N)
    assert O.rel_l2(rt, grid) < 1e-12


@pytest.mark.parametrize("name", golden_names())
def test_plan_matches_reference_fixture(name):
    """fixtures = outputs of the unmodified reference (tests/golden/make_golden.py)"""
    _torch()
    g = load_golden(name)
    grid = O.grid_values(g["seed"], *g["N"])
    got, _, _ = gpu_forward(grid, g["p"], g["custom"], g["is_oned"], g["is_equalxy"])
    assert got[0].params == g["params"]
    assert O.rel_l2(O.gather_output(got), O.gather_output(g["boxes"])) < 1e-12


@pytest.mark.parametrize("N,p,oned,custom", [((64, 32, 128), 1, 0, {P.P1: 1}), ((32, 64, 64), 4, 0, {P.P1: 2}),
                                             ((64, 64, 64), 4, 1, {P.P1: 4, P.S: 1})])
def test_single_precision(oracle, N, p, oned, custom):
    _torch()
    grid = O.grid_values(9, *N)
    want = O.gather_output(oracle.execute(grid, p, oracle.resolve_params(*N, p, custom), oned, 0))
    got, _, back = gpu_forward(grid, p, custom, oned, 0, bits=32, inverse_too=True)
    assert O.rel_l2(O.gather_output(got, np.complex64), want) < 1e-5
    assert O.rel_l2(gather_input(got, back, np.complex64) / np.prod(N), grid) < 1e-5


def test_host_arrays_and_ramp_known_answer():
    """the reference driver's own usage: calloc'd host array, ramp input, read (0,0,z<4) through
    ostride (run-fft.c:46-61, 452-503)"""
    _torch()
    N = (64, 64, 64)
    got, _, _ = gpu_forward(O.ramp_values(*N), 1, {P.P1: 1}, host_arrays=True)
    b = got[0]
    vals = [b.data[z * b.ostride[2]] for z in range(4)]
    n = 64
    want = [n ** 3 * (n - 1) / 2 * 111 + 0j] + [n ** 3 / (np.exp(-2j * np.pi * k / n) - 1) for k in (1, 2, 3)]
    np.testing.assert_allclose(vals, want, rtol=1e-12)


def test_full_size_properties():
    """BASELINE config 2 size (512^3 complex128, one GPU): size-independent checks -
    a plane wave gives one spike, Parseval, and forward->backward returns N*x."""
    torch = _torch()
    import offt_b200 as ob
    n = 512
    with local_world(1):
        plan = ob.Plan(n, n, n, is_notest=1, custom={P.P1: 1, P.S: 1})
        x = torch.arange(n, device="cuda", dtype=torch.float64)
        ph = 2 * np.pi * (3 * x[:, None, None] + 5 * x[None, :, None] + 7 * x[None, None, :]) / n
        a = torch.polar(torch.ones_like(ph), ph).contiguous()
        del ph
        plan.execute(a)
        assert abs(a[3, 5, 7].item() - n ** 3) < 1e-3
        a[3, 5, 7] = 0
        assert a.abs().max().item() < 1e-4
        g = torch.Generator(device="cuda"); g.manual_seed(1)
        a = torch.view_as_complex(torch.rand((n, n, n, 2), device="cuda", dtype=torch.float64, generator=g) * 2 - 1).contiguous()
        ref = a.clone()
        e_in = (ref.abs() ** 2).sum().item()
        plan.execute(a)
        e_out = (a.abs() ** 2).sum().item()
        assert abs(e_out / (n ** 3) - e_in) / e_in < 1e-12
        plan.execute_inverse(a)
        a /= n ** 3
        err = (torch.linalg.vector_norm(a - ref) / torch.linalg.vector_norm(ref)).item()
        assert err < 1e-12
        plan.fin()


def test_unsupported_inputs_fail_loudly():
    _torch()
    import offt_b200 as ob
    with local_world(1):
        with pytest.raises(ob.OfftError):
            ob.Plan(16, 16, 7001, custom={P.P1: 1})         # not a power of two and too long for the shared-memory kernel
        with pytest.raises(ob.OfftError):
            ob.Plan(16, 16, 16384, custom={P.P1: 1})        # power of two beyond the kernels
    with local_world(4):
        with pytest.raises(ob.OfftError):
            ob.Plan(2, 16, 16, custom={P.P1: 4}, rank=0)    # P1 outside the reference's range: a rank without an x plane


# ---------------------------------------------------------------- any length, uneven splits (fft_generic.cu)
@pytest.mark.parametrize("bits", [64, 32])
@pytest.mark.parametrize("n", [1, 3, 5, 6, 7, 9, 10, 12, 15, 18, 20, 49, 60, 100, 121, 127, 210, 243, 360, 1000, 1001, 1536, 2187, 3000])
def test_rows_any_length(oracle, n, bits):
    """lengths with odd factors and primes, contiguous and strided, forward and backward (the reference hands any
    length to FFTW, offt-compute.c:338, 416, 443)"""
    torch = _torch()
    import offt_b200 as ob
    rng = np.random.default_rng(n)
    rows = 12
    a = (rng.uniform(-1, 1, (rows, n)) + 1j * rng.uniform(-1, 1, (rows, n)))
    want = np.fft.fft(a, axis=1)
    chk = a[:2].copy()
    oracle.dft_rows(chk, n, 1, n, 2, -1)        # the oracle's own 1-D arithmetic agrees with numpy on this length
    assert O.rel_l2(chk, want[:2]) < 1e-12
    cdt = torch.complex128 if bits == 64 else torch.complex64
    with local_world(1):
        d = torch.from_numpy(a).to("cuda").to(cdt).contiguous()
        ob.fft_rows(d, n, 1, n, rows, sign=-1, bits=bits)
        assert O.rel_l2(d.cpu().numpy(), want) < TOL[bits]
        d = torch.from_numpy(np.ascontiguousarray(a.T)).to("cuda").to(cdt).contiguous()
        ob.fft_rows(d, n, rows, 1, rows, sign=-1, bits=bits)
        assert O.rel_l2(d.cpu().numpy().T, want) < TOL[bits]
        ob.fft_rows(d, n, rows, 1, rows, sign=+1, bits=bits)
        assert O.rel_l2(d.cpu().numpy().T / n, a) < TOL[bits]


UNEVEN_CASES = [
    # N, p, is_oned, is_equalxy, custom: non-power-of-two grids, uneven process grids, exact-count bits of _V_
    ((12, 10, 9), 1, 0, 0, {P.P1: 1}),
    ((12, 10, 9), 1, 0, 0, {P.P1: 1, P.S: 1}),
    ((12, 10, 9), 4, 0, 0, {P.P1: 2, P.V: 3}),
    ((12, 10, 9), 4, 0, 0, {P.P1: 2, P.V: 0, P.S: 1}),
    ((12, 10, 9), 3, 1, 0, {P.P1: 3, P.V: 3, P.T2: 2}),
    ((12, 10, 9), 3, 1, 0, {P.P1: 1, P.V: 1, P.T1: 5}),
    ((20, 12, 18), 6, 0, 0, {P.P1: 3, P.V: 3, P.T1: 3, P.T2: 4}),
    ((20, 12, 18), 6, 0, 0, {P.P1: 2, P.V: 2, P.S: 1, P.RY: 3}),
    ((16, 16, 16), 3, 0, 0, {P.P1: 3}),                         # power-of-two lengths, uneven split
    ((16, 32, 16), 6, 0, 0, {P.P1: 2, P.T1: 3, P.T2: 2}),
    ((18, 18, 6), 4, 0, 1, {P.P1: 2}),                          # equalxy with odd-factor lengths
    ((30, 42, 70), 5, 1, 0, {P.P1: 5, P.S: 1}),
    ((30, 42, 70), 7, 1, 0, {P.P1: 1}),
    ((96, 80, 48), 8, 0, 0, {P.P1: 4, P.T1: 5, P.T2: 7, P.W1: 2, P.W2: 1}),
    ((27, 20, 45), 8, 1, 0, {P.P1: 1, P.S: 1, P.T1: 4}),        # 8 ranks: shares of 2-3 rows and 5-6 columns
    ((27, 20, 45), 8, 1, 0, {P.P1: 1, P.S: 1, P.T1: 4, P.W1: 0}),   # in place, plane strides 24*M3 vs 20*M3: the backward tile order matters
    ((27, 20, 45), 8, 1, 0, {P.P1: 8, P.T2: 2}),
    ((27, 20, 45), 8, 0, 0, {P.P1: 2, P.S: 1}),
    ((27, 11, 45), 2, 1, 0, {P.P1: 1, P.S: 1, P.T1: 2, P.W1: 0}),   # the same skew (12 rows against 11) at p = 2 and in a pencil
    ((12, 11, 15), 4, 0, 0, {P.P1: 2, P.S: 1, P.T1: 2, P.W1: 0, P.T2: 4, P.W2: 0}),
    ((12, 13, 15), 8, 0, 0, {P.P1: 2, P.S: 1, P.T1: 2, P.W1: 0, P.T2: 2, P.W2: 0}),   # a pencil grid with the skew: 16 rows against 14
]


@pytest.mark.parametrize("N,p,oned,eq,custom", UNEVEN_CASES)
def test_uneven_and_any_length_plans_match_oracle(oracle, N, p, oned, eq, custom):
    _torch()
    grid = O.grid_values(6, *N)
    v = oracle.resolve_params(*N, p, custom)
    want = oracle.execute(grid, p, v, oned, eq)
    got, launches, back = gpu_forward(grid, p, custom, oned, eq, inverse_too=True)
    assert launches > 0
    for a, b in zip(got, want):
        assert a.params == b.params
        assert (a.istart, a.isize, a.istride, a.ostart, a.osize, a.ostride, a.alloc) == \
               (b.istart, b.isize, b.istride, b.ostart, b.osize, b.ostride, b.alloc)
    A, B = O.gather_output(got), O.gather_output(want)
    assert not np.isnan(A).any()
    assert O.rel_l2(A, B) < 1e-12 and O.rel_l2(A, np.fft.fftn(grid)) < 1e-12
    assert O.rel_l2(gather_input(got, back) / np.prod(N), grid) < 1e-12


@pytest.mark.parametrize("N,p,oned,eq,custom", [CASES[i] for i in (0, 4, 9, 11, 15, 16, 19)])
def test_generic_kernel_on_power_of_two_plans(oracle, N, p, oned, eq, custom):
    """the any-length kernel forced onto plans the fast kernels normally take: same results"""
    _torch()
    import offt_b200 as ob
    grid = O.grid_values(5, *N)
    want = O.gather_output(oracle.execute(grid, p, oracle.resolve_params(*N, p, custom), oned, eq))
    ob.set_force_generic(True)
    try:
        got, launches, back = gpu_forward(grid, p, custom, oned, eq, inverse_too=True)
    finally:
        ob.set_force_generic(False)
    assert launches > 0
    assert O.rel_l2(O.gather_output(got), want) < 1e-12
    assert O.rel_l2(gather_input(got, back) / np.prod(N), grid) < 1e-12


def test_uneven_single_precision(oracle):
    _torch()
    N, p, custom = (20, 12, 18), 6, {P.P1: 3, P.V: 3}
    grid = O.grid_values(9, *N)
    want = O.gather_output(oracle.execute(grid, p, oracle.resolve_params(*N, p, custom), 0, 0))
    got, _, back = gpu_forward(grid, p, custom, 0, 0, bits=32, inverse_too=True)
    assert O.rel_l2(O.gather_output(got, np.complex64), want) < 1e-5
    assert O.rel_l2(gather_input(got, back, np.complex64) / np.prod(N), grid) < 1e-5


def test_async_execution_on_the_callers_stream(oracle):
    """offtb_plan_set_stream / offtb_plan_set_async: the transform is enqueued on the caller's CUDA stream and
    offt_3d_execute returns at once; work enqueued before and after it on that stream is ordered around it."""
    torch = _torch()
    import offt_b200 as ob
    N = (64, 128, 256)
    grid = O.grid_values(21, *N)
    want = np.fft.fftn(grid)
    with local_world(1):
        plan = ob.Plan(*N, is_notest=1, custom={P.P1: 1, P.S: 1})
        st = torch.cuda.Stream()
        plan.set_stream(st.cuda_stream)
        plan.set_async(True)
        host = torch.from_numpy(grid.reshape(-1).copy()).pin_memory()
        dev = torch.empty_like(host, device="cuda")
        out = torch.empty_like(host)
        with torch.cuda.stream(st):
            dev.copy_(host, non_blocking=True)      # H2D, the transform and D2H all ride the same stream
            plan.execute(dev)
            dev.mul_(2.0)
            out.copy_(dev, non_blocking=True)
        st.synchronize()
        assert O.rel_l2(out.numpy().reshape(N) / 2.0, want) < 1e-12
        plan.set_async(False)
        plan.set_stream(0)
        plan.fin()


@pytest.mark.parametrize("N,bits", [((32, 128, 64), 64), ((128, 16, 256), 32), ((8, 8, 8), 64)])
def test_single_rank_xyz_schedule(oracle, N, bits):
    """one rank, _S_ = 1: the schedule that swaps the outer axes in its z pass (plan.cu, L_fftz_swap) - anisotropic
    grids, both precisions, forward against the oracle and numpy, then backward"""
    _torch()
    grid = O.grid_values(17, *N)
    custom = {P.P1: 1, P.S: 1}
    want = O.gather_output(oracle.execute(grid, 1, oracle.resolve_params(*N, 1, custom), 0, 0))
    got, launches, back = gpu_forward(grid, 1, custom, bits=bits, inverse_too=True)
    assert launches == 3
    cdt = np.complex128 if bits == 64 else np.complex64
    A = O.gather_output(got, cdt)
    assert O.rel_l2(A, want) < TOL[bits] and O.rel_l2(A, np.fft.fftn(grid)) < TOL[bits]
    assert O.rel_l2(gather_input(got, back, cdt) / np.prod(N), grid) < TOL[bits]


# ---------------------------------------------------------------- real-to-complex plans (is_r2c, offt-compute.c:63, 334-336, 960-961)
@pytest.mark.parametrize("name", golden_names("r2c"))
def test_r2c_plan_matches_reference_fixture(name):
    """fixtures = outputs of the unmodified reference run with is_r2c = 1 (tests/golden/make_golden.py)"""
    _torch()
    g = load_golden(name, "r2c")
    grid = O.grid_values(g["seed"], *g["N"])
    got, _, _ = gpu_forward(grid, g["p"], g["custom"], g["is_oned"], g["is_equalxy"], is_r2c=1)
    assert got[0].params == g["params"]
    for a, b in zip(got, g["boxes"]):
        assert (a.istart, a.isize, a.istride, a.ostart, a.osize, a.ostride, a.alloc) == \
               (b.istart, b.isize, b.istride, b.ostart, b.osize, b.ostride, b.alloc)
    assert O.rel_l2(O.gather_output_r2c(got), O.gather_output_r2c(g["boxes"])) < 1e-12


@pytest.mark.parametrize("N,p,oned,custom,bits", [
    ((64, 32, 128), 1, 0, {P.P1: 1}, 64), ((64, 32, 128), 1, 0, {P.P1: 1, P.S: 1}, 64), ((32, 64, 256), 4, 1, {P.P1: 4}, 64),
    ((32, 64, 250), 4, 1, {P.P1: 1, P.S: 1, P.T1: 5}, 64), ((30, 24, 50), 6, 0, {P.P1: 3, P.V: 3}, 64), ((64, 64, 64), 8, 0, {P.P1: 2, P.T1: 8, P.T2: 5}, 64),
    ((64, 32, 128), 4, 0, {P.P1: 2}, 32),
    # in place with unequal plane strides around phase 1 and no slot in flight: forward tiles go up, backward tiles down
    ((27, 11, 50), 2, 1, {P.P1: 1, P.S: 1, P.T1: 2, P.W1: 0}, 64), ((12, 11, 30), 4, 0, {P.P1: 2, P.S: 1, P.T1: 2, P.W1: 0, P.T2: 4, P.W2: 0}, 64)])
def test_r2c_matches_numpy_and_round_trips(N, p, oned, custom, bits):
    """half spectrum against numpy.fft.rfftn; the backward transform (complex-to-real) returns N * the real input"""
    _torch()
    grid = O.grid_values(12, *N)
    want = np.fft.rfftn(grid.real)
    got, launches, back = gpu_forward(grid, p, custom, oned, 0, bits=bits, inverse_too=True, is_r2c=1)
    assert launches > 0
    cdt = np.complex128 if bits == 64 else np.complex64
    assert O.rel_l2(O.gather_output_r2c(got, cdt), want) < TOL[bits]
    if bits == 32:
        back = [b.view(np.float32).astype(np.float64).view(np.complex128) for b in back]
    assert O.rel_l2(O.gather_input_r2c(got, back) / np.prod(N), grid.real) < TOL[bits]


def test_timer_buckets_are_filled_by_default():
    """po->t[] (offt.h:171-188), what run-fft.c prints per repetition (run-fft.c:399-413): ALL and the four FFT
    buckets carry device time after every execute without any opt-in; the fused-away steps stay zero"""
    torch = _torch()
    import offt_b200 as ob
    ALL, INIT1, WAIT1, TEST1, INIT2, WAIT2, TEST2, FFTz, FFTy1, FFTy2, FFTx, TRANSPOSE, PACK1, UNPACK1, PACK2, UNPACK2 = range(16)
    N = (64, 64, 64)
    with local_world(1):
        plan = ob.Plan(*N, is_notest=1, custom={P.P1: 1})
        a = torch.zeros(plan.alloc_elems, dtype=torch.complex128, device="cuda")
        plan.execute(a)
        t = plan.t
        assert t[ALL] > 0 and t[FFTz] > 0 and t[FFTy1] > 0 and t[FFTx] > 0
        assert t[FFTz] + t[FFTy1] + t[FFTy2] + t[FFTx] <= t[ALL] * 1.05
        assert t[TRANSPOSE] == 0 and t[PACK1] == 0 and t[TEST1] == 0
        plan.fin()
    with local_world(4):
        plans = [ob.Plan(*N, is_notest=1, custom={P.P1: 2}, rank=r) for r in range(4)]
        arrs = [torch.zeros(pl.alloc_elems, dtype=torch.complex128, device="cuda") for pl in plans]
        ob.execute_group(plans, arrs)
        t = plans[0].t
        assert all(t[k] > 0 for k in (ALL, FFTz, FFTy1, FFTy2, FFTx, INIT1, INIT2))
        for pl in plans:
            pl.fin()


@pytest.mark.parametrize("S", [1, 0])
def test_full_size_512_cubed_against_an_independent_fft(S):
    """BASELINE configs[1] at its full size, element by element: 512^3 complex128 on one GPU against scipy's pocketfft
    (numpy.fft.fftn where scipy is absent) on the same seeded grid, read through ostart/osize/ostride"""
    torch = _torch()
    import offt_b200 as ob
    from offt_b200 import layout
    try:
        import scipy.fft
        fftn = lambda a: scipy.fft.fftn(a, workers=-1)   # noqa: E731
    except ImportError:
        fftn = np.fft.fftn
    n = 512
    rng = np.random.default_rng(512)
    grid = rng.uniform(-1, 1, (n, n, n)) + 1j * rng.uniform(-1, 1, (n, n, n))
    with local_world(1):
        plan = ob.Plan(n, n, n, is_notest=1, custom={P.P1: 1, P.S: S})
        box = plan.box()
        a = torch.from_numpy(layout.scatter_input(box, grid, plan.alloc_elems)).to("cuda")
        plan.execute(a)
        got = layout.gather_output([box], [a.cpu().numpy()], (n, n, n))
        plan.fin()
    want = fftn(grid)
    del grid
    num = np.linalg.norm((got - want).ravel())
    den = np.linalg.norm(want.ravel())
    assert num / den < 1e-12
