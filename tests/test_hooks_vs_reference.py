"""CPU: the tunable-parameter hooks of the library against the UNMODIFIED reference functions of the same names
(oracle/_ref/ref_hooks links offt-compute.o / offt-tuning.o compiled from /root/reference): value grids, default
heuristics, grid_value_floor/ceil, the index -> value conversion with its ADJUST_POINT repairs, and the feasibility
verdict (including WHICH tunable is reported) on hundreds of random points of the search space per configuration."""
import ctypes as C
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
HOOKS = ROOT / "oracle" / "_ref" / "ref_hooks"
sys.path.insert(0, str(ROOT))

CONFIGS = [  # Nx Ny Nz p is_oned is_W0 is_notest
    (64, 32, 128, 4, 0, 0, 1), (256, 256, 256, 4, 1, 0, 0), (1024, 1024, 1024, 8, 1, 0, 1), (2048, 2048, 2048, 8, 0, 0, 1),
    (2048, 1024, 512, 8, 1, 1, 1), (512, 512, 512, 1, 0, 0, 1), (512, 512, 512, 64, 0, 0, 0), (16, 8, 32, 2, 1, 0, 1),
]


def _reference(cfg, npoints, seed):
    out = subprocess.run([str(HOOKS)] + [str(v) for v in cfg] + [str(npoints), str(seed)], capture_output=True, text=True, timeout=120, check=True).stdout
    ranges, default, points, fc = {}, None, [], []
    for line in out.splitlines():
        f = line.split()
        if f[0] == "range":
            ranges[int(f[1])] = [int(x) for x in f[3:]]
        elif f[0] == "default":
            default = [int(x) for x in f[1:]]
        elif f[0] == "point":
            arrow, inf = f.index("->"), f.index("infeasible")
            points.append(([int(x) for x in f[1:arrow]], [int(x) for x in f[arrow + 1:inf]], int(f[inf + 1]), int(f[inf + 2])))
        elif f[0] == "floorceil":
            fc.append(tuple(int(x) for x in f[1:]))
    return ranges, default, points, fc


def _reference_printers(cfg):
    out = subprocess.run([str(HOOKS)] + [str(v) for v in cfg] + ["0", "1"], capture_output=True, text=True, timeout=120, check=True).stdout
    lines = {l.split(": ", 1)[0]: l.split(": ", 1)[1] for l in out.splitlines() if l.startswith(("print_params: ", "offt_print_time: "))}
    return lines["print_params"], lines["offt_print_time"]


def _capture_c_stdout(fn):
    """what a C function of the library writes to stdout"""
    import os
    import tempfile
    libc = C.CDLL(None)
    sys.stdout.flush()
    with tempfile.TemporaryFile() as tmp:
        saved = os.dup(1)
        os.dup2(tmp.fileno(), 1)
        try:
            fn()
            libc.fflush(None)
        finally:
            os.dup2(saved, 1)
            os.close(saved)
        tmp.seek(0)
        return tmp.read().decode()


@pytest.mark.skipif(not HOOKS.exists(), reason="oracle/_ref/ref_hooks not built (needs /root/reference)")
def test_printers_match_the_reference_byte_for_byte():
    """print_params and offt_print_time (offt.h:243-244): the lines run-fft.c's users read and scripts parse"""
    from offt_b200.binding import GES, PARAM_COUNT, lib
    cfg = (1024, 1024, 1024, 8, 1, 0, 1)
    want_params, want_time = _reference_printers(cfg)
    import offt_b200 as ob
    v = (C.c_int * PARAM_COUNT)(*ob.params_default(*cfg[:4], cfg[5], cfg[6]))
    t = (C.c_double * GES)(*[0.001 * (i + 1) + 0.0000049 * i for i in range(GES)])
    lib.print_params.argtypes = [C.POINTER(C.c_int)]
    lib.print_params.restype = None
    lib.offt_print_time.argtypes = [C.POINTER(C.c_double)]
    lib.offt_print_time.restype = None
    assert _capture_c_stdout(lambda: lib.print_params(v)).rstrip("\n") == want_params
    assert _capture_c_stdout(lambda: lib.offt_print_time(t)).rstrip("\n") == want_time


@pytest.mark.skipif(not HOOKS.exists(), reason="oracle/_ref/ref_hooks not built (needs /root/reference)")
@pytest.mark.parametrize("cfg", CONFIGS)
def test_hooks_match_the_reference_functions(cfg):
    import offt_b200 as ob
    from offt_b200.binding import OfftParams, OfftPlan, PARAM_COUNT, lib
    Nx, Ny, Nz, p, oned, W0, notest = cfg
    ranges, default, points, fc = _reference(cfg, 400, 12345)
    po = OfftPlan()
    par = OfftParams()
    po.Nx, po.Ny, po.Nz, po.p, po.is_oned, po.is_W0, po.is_notest = Nx, Ny, Nz, p, oned, W0, notest
    po.params = C.pointer(par)
    # params_range_setup
    lists = (C.POINTER(C.c_int) * PARAM_COUNT)()
    sizes = (C.c_int * PARAM_COUNT)()
    lib.params_range_setup.argtypes = [C.POINTER(OfftPlan), C.POINTER(C.POINTER(C.c_int)), C.POINTER(C.c_int)]
    lib.params_range_setup.restype = None
    lib.params_range_setup(C.byref(po), lists, sizes)
    for i in range(PARAM_COUNT):
        assert [lists[i][k] for k in range(sizes[i])] == ranges[i], f"value grid of tunable {i}"
    # params_set_default
    lib.params_set_default.argtypes = [C.POINTER(OfftPlan)]
    lib.params_set_default.restype = None
    lib.params_set_default(C.byref(po))
    assert list(par.v) == default
    # params_convert (backward) + is_infeasible_point
    lib.params_convert.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_long), C.POINTER(OfftPlan), C.POINTER(C.POINTER(C.c_int)), C.POINTER(C.c_int)]
    lib.params_convert.restype = None
    lib.is_infeasible_point.argtypes = [C.POINTER(OfftPlan), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    for idx, values, inf, bad in points:
        ahv = (C.c_long * PARAM_COUNT)(*idx)
        v = (C.c_int * PARAM_COUNT)()
        lib.params_convert(1, v, ahv, C.byref(po), lists, sizes)
        assert list(v) == values, f"params_convert at indices {idx}"
        b = C.c_int(-7)
        r = lib.is_infeasible_point(C.byref(po), v, C.byref(b))
        assert (r, b.value) == (inf, bad), f"is_infeasible_point at {values}: got {(r, b.value)}, reference {(inf, bad)}"
        # forward conversion returns the indices of the (repaired) values when they are still on the grid
        if all(values[i] in ranges[i] for i in range(PARAM_COUNT)):
            back = (C.c_long * PARAM_COUNT)()
            lib.params_convert(0, v, back, C.byref(po), lists, sizes)
            assert [ranges[i][back[i]] for i in range(PARAM_COUNT)] == values
    # grid_value_floor / grid_value_ceil
    for fn in (lib.grid_value_floor, lib.grid_value_ceil):
        fn.argtypes = [C.c_int, C.POINTER(C.POINTER(C.c_int)), C.POINTER(C.c_int), C.c_int, C.c_int]
    for i, raw, fv, cv, fi, ci in fc:
        got = (lib.grid_value_floor(0, lists, sizes, i, raw), lib.grid_value_ceil(0, lists, sizes, i, raw),
               lib.grid_value_floor(1, lists, sizes, i, raw), lib.grid_value_ceil(1, lists, sizes, i, raw))
        assert got == (fv, cv, fi, ci), f"grid_value_floor/ceil of tunable {i} at {raw}"
    assert ob  # the binding loaded the product library


SIMPLEX_CONFIGS = [  # (Nx Ny Nz p is_oned is_W0 is_notest), tuning_mode, is_r2c, seed
    ((64, 32, 128, 4, 0, 0, 1), 0, 0, 7), ((256, 256, 256, 4, 1, 0, 0), 0, 0, 11), ((1024, 1024, 1024, 8, 1, 0, 1), 2, 0, 3),
    ((2048, 2048, 2048, 8, 0, 0, 1), 0, 0, 5), ((2048, 1024, 512, 8, 1, 1, 1), 0, 0, 9), ((2048, 1024, 512, 8, 1, 0, 1), 1, 0, 13),
    ((512, 512, 512, 64, 0, 0, 0), 0, 0, 2), ((64, 32, 130, 4, 0, 0, 1), 0, 1, 21), ((16, 8, 32, 2, 1, 0, 1), 0, 0, 1),
]


@pytest.mark.skipif(not HOOKS.exists(), reason="oracle/_ref/ref_hooks not built (needs /root/reference)")
@pytest.mark.parametrize("cfg,mode,r2c,seed", SIMPLEX_CONFIGS)
def test_initial_simplex_matches_reference(cfg, mode, r2c, seed, tmp_path):
    """write_initial_simplex (offt-tuning.c:426-737): with the same srand() seed the library draws the same 25 vertices,
    coordinate by coordinate, as the unmodified reference function - for every tuning mode, r2c, is_W0 and is_notest"""
    import offt_b200 as ob
    out = subprocess.run([str(HOOKS)] + [str(v) for v in cfg] + ["0", "1", str(seed), str(mode), str(r2c)],
                         capture_output=True, text=True, timeout=120, check=True).stdout
    want = [[int(x) for x in line.split()[2:]] for line in out.splitlines() if line.startswith("simplex")]
    assert len(want) == 25 and all(len(w) == 24 for w in want)
    po = ob.binding.OfftPlan()
    po.Nx, po.Ny, po.Nz, po.p, po.is_oned, po.is_W0, po.is_notest = cfg
    po.tuning_mode, po.is_r2c = mode, r2c
    po.user_vertex_file = str(tmp_path / "uv").encode()
    lib = ob.lib
    v_list = (C.POINTER(C.c_int) * 24)()
    v_size = (C.c_int * 24)()
    lib.params_range_setup.argtypes = [C.POINTER(ob.binding.OfftPlan), C.POINTER(C.POINTER(C.c_int)), C.POINTER(C.c_int)]
    lib.params_range_setup.restype = None
    lib.write_initial_simplex.argtypes = lib.params_range_setup.argtypes
    lib.write_initial_simplex.restype = None
    lib.params_range_setup(C.byref(po), v_list, v_size)
    C.CDLL(None).srand(seed)
    lib.write_initial_simplex(C.byref(po), v_list, v_size)
    got = [[int(x) for x in line.split()] for line in (tmp_path / "uv").read_text().splitlines()]
    assert got == want
