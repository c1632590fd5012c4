"""ctypes binding of libofft_b200.so: the reference's plan / execute / destroy API
(include/offt.h, mirroring offt.h:235-244 of rchyena/offt) plus include/offt_b200.h."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

HERE = Path(__file__).resolve().parent
# OFFTB_LIB selects a variant build of the same library (tools/variants.sh); the default is the product
LIB_PATH = Path(os.environ.get("OFFTB_LIB") or HERE / "lib" / "libofft_b200.so")
PARAM_COUNT = 24
GES = 16
PARAM_NAMES = ("P1 T1 W1 Px1 Py1 Fz FP1 Ux1 Uz1 FU1 Fy1 Ry T2 W2 Pz2 Px2 Fy2 FP2 Uz2 Uy2 FU2 Fx V S").split()


class P:
    """parameter indices, offt.h:74-98"""
    P1, T1, W1, Px1, Py1, Fz, FP1, Ux1, Uz1, FU1, Fy1, Ry, T2, W2, Pz2, Px2, Fy2, FP2, Uz2, Uy2, FU2, Fx, V, S = range(24)


class OfftError(RuntimeError):
    pass


class OfftParams(C.Structure):
    _fields_ = [("is_converged", C.c_int), ("is_infeasible", C.c_int), ("is_in_database", C.c_int),
                ("v", C.c_int * PARAM_COUNT)]


class OfftComm(C.Structure):
    _fields_ = [("p1", C.c_int), ("p2", C.c_int), ("comm1", C.c_void_p), ("comm2", C.c_void_p),
                ("group1", C.c_void_p), ("group2", C.c_void_p)] + \
               [(n, C.c_int) for n in "M1 M2 M3 M4 F1 F2 F3 F4 m1 m2 m3 m4 b1 b2 b3 b4".split()] + \
               [(n, C.c_int * 3) for n in "istart isize istride ostart osize ostride".split()]


class OfftPlan(C.Structure):
    # the library is built without SHSONG_HOPPER / SHSONG_EDISON, so bad_tile_count is present (include/offt.h)
    _fields_ = [(n, C.c_int) for n in ("p rank Nx Ny Nz is_r2c fftw_flag ah_strategy max_loop tuning_mode is_W0 "
                                       "extrapolation_window is_oned is_a2a is_equalxy is_notest").split()] + \
               [("t_init", C.c_double * 4), ("t", C.c_double * GES), ("bad_tile_count", C.c_int),
                ("point_database_file", C.c_char * 256), ("user_vertex_file", C.c_char * 256),
                ("params", C.POINTER(OfftParams)), ("comm", C.POINTER(OfftComm)),
                ("buffer_chunk", C.c_void_p), ("buffers1", C.c_void_p), ("buffers2", C.c_void_p),
                ("pt_transpose", C.c_void_p), ("pt_transpose_list", C.c_void_p), ("pt_transpose_list_size", C.c_int),
                ("p1d_x", C.c_void_p), ("p1d_y", C.c_void_p), ("p1d_z", C.c_void_p), ("p1d_x_t", C.c_void_p),
                ("p1d_y_t", C.c_void_p), ("p1d_x_s_list", C.c_void_p), ("p1d_y_s_list", C.c_void_p),
                ("p1d_xy_s_list_size", C.c_int), ("b200", C.c_void_p)]


def _load():
    if not LIB_PATH.exists():
        raise OfftError(f"{LIB_PATH} is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback)")
    L = C.CDLL(str(LIB_PATH))
    ll, i, vp, d = C.c_longlong, C.c_int, C.c_void_p, C.c_double
    L.offtb_last_error.restype = C.c_char_p
    L.offtb_clear_error.restype = None
    L.offt_3d_init.restype = C.POINTER(OfftPlan)
    L.offt_3d_init.argtypes = [i, i, i, vp, vp] + [i] * 11 + [C.POINTER(OfftParams)]
    L.offt_3d_execute.restype = None
    L.offt_3d_execute.argtypes = [C.POINTER(OfftPlan), vp, vp, i]
    L.offt_3d_execute_inverse.argtypes = [C.POINTER(OfftPlan), vp, vp]
    L.offt_3d_fin.restype = None
    L.offt_3d_fin.argtypes = [C.POINTER(OfftPlan)]
    L.offtb_execute_group.argtypes = [C.POINTER(C.POINTER(OfftPlan)), C.POINTER(vp), i, i]
    L.offtb_world_init.argtypes = [i, i, i, vp]
    L.offtb_world_init_local.argtypes = [i, i]
    L.offtb_get_unique_id.argtypes = [vp]
    L.offtb_world_fin.restype = None
    L.offtb_plan_set_stream.argtypes = [C.POINTER(OfftPlan), vp]
    L.offtb_plan_set_async.argtypes = [C.POINTER(OfftPlan), i]
    L.offtb_plan_set_stage_timing.argtypes = [C.POINTER(OfftPlan), i]
    L.offtb_plan_alloc_elems.restype = ll
    L.offtb_plan_alloc_elems.argtypes = [C.POINTER(OfftPlan)]
    L.offtb_alloc_elems.restype = ll
    L.offtb_alloc_elems.argtypes = [i] * 5
    L.offtb_plan_last_ms.restype = d
    L.offtb_plan_last_ms.argtypes = [C.POINTER(OfftPlan)]
    L.offtb_plan_last_launches.argtypes = [C.POINTER(OfftPlan)]
    L.offtb_plan_precision.argtypes = [C.POINTER(OfftPlan)]
    L.offtb_plan_stage_ms.argtypes = [C.POINTER(OfftPlan), C.POINTER(d), i]
    L.offtb_tile_visited.restype = i
    L.offtb_tile_visited.argtypes = [i, i, i, i]
    L.offtb_exchange_block_elems.restype = ll
    L.offtb_exchange_block_elems.argtypes = [C.POINTER(OfftPlan), i, i]
    L.offtb_fft_rows.restype = d
    L.offtb_fft_rows.argtypes = [vp, i, ll, ll, ll, i, i, i, vp]
    L.offtb_fft_launch_raw.restype = d
    L.offtb_fft_launch_raw.argtypes = [vp, vp, i, i, i, ll, C.POINTER(ll), C.POINTER(ll)] + [i] * 8 + [vp]
    L.offtb_params_default.restype = None
    L.offtb_params_default.argtypes = [i] * 6 + [C.POINTER(i)]
    L.offtb_is_infeasible_point.argtypes = [i] * 4 + [C.POINTER(i), C.POINTER(i)]
    L.offtb_params_adjust.restype = None
    L.offtb_params_adjust.argtypes = [i] * 5 + [C.POINTER(i)]
    L.offtb_params_range.restype = None
    L.offtb_params_range.argtypes = [i] * 4 + [C.POINTER(i), i, C.POINTER(i)]
    L.offtb_comm_fill.argtypes = [C.POINTER(OfftComm)] + [i] * 8
    L.offtb_comm_fill_r2c.argtypes = [C.POINTER(OfftComm)] + [i] * 9
    L.offtb_alloc_elems_r2c.restype = ll
    L.offtb_alloc_elems_r2c.argtypes = [i] * 6
    L.offtb_check_supported.argtypes = [i] * 5
    L.offtb_tune.argtypes = [C.POINTER(OfftPlan), vp, vp, i, i]
    L.offtb_tune_ex.argtypes = [C.POINTER(OfftPlan), vp, vp, i, i, i, i]
    L.offtb_tune_harmony.argtypes = [C.POINTER(OfftPlan), i, i, i, i]
    L.offtb_set_exit_on_error(0)   # Python raises instead of exit(-1)
    return L


lib = _load()


def _err() -> str:
    return (lib.offtb_last_error() or b"").decode()


def _check(rc, what):
    if rc != 0:
        raise OfftError(f"{what}: {_err()}")


def get_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _check(lib.offtb_get_unique_id(buf), "offtb_get_unique_id")
    return buf.raw


def world_init(rank: int, size: int, device: int = -1, unique_id: bytes | None = None):
    buf = C.create_string_buffer(unique_id, 128) if unique_id else None
    _check(lib.offtb_world_init(rank, size, device, buf), "offtb_world_init")


def world_init_local(size: int, device: int = 0):
    _check(lib.offtb_world_init_local(size, device), "offtb_world_init_local")


def world_fin():
    lib.offtb_world_fin()


def world_size() -> int:
    return lib.offtb_world_size()


def world_rank() -> int:
    return lib.offtb_world_rank()


def set_force_generic(on: bool):
    _check(lib.offtb_set_force_generic(int(on)), "offtb_set_force_generic")


def set_default_precision(bits: int):
    _check(lib.offtb_set_default_precision(bits), "offtb_set_default_precision")


def _ptr(a) -> int:
    """address of a torch tensor (host or device) or a numpy array"""
    if a is None:
        return 0
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    if hasattr(a, "ctypes"):
        return a.ctypes.data
    return int(a)


class Plan:
    """offt_3d_init / offt_3d_execute / offt_3d_fin with the reference's argument list
    (offt.h:235-241).  `custom` maps parameter index -> value like run-fft.c's -d/-T/-W/... flags
    (unset entries are -1 = "use the default", run-fft.c:166-167)."""

    def __init__(self, Nx, Ny, Nz, *, is_oned=0, is_equalxy=0, is_a2a=0, is_notest=0, is_W0=0, max_loop=0,
                 is_r2c=0, custom=None, rank=None, array=None):
        if rank is not None:
            _check(lib.offtb_world_set_rank(rank), "offtb_world_set_rank")
        cp = OfftParams()
        for k in range(PARAM_COUNT):
            cp.v[k] = -1
        for k, v in (custom or {}).items():
            cp.v[k] = v
        self._custom = cp
        ptr = _ptr(array)
        self.po = lib.offt_3d_init(Nx, Ny, Nz, ptr, ptr, is_r2c, 0, is_oned, is_a2a, is_equalxy, is_notest, 0,
                                   max_loop, 0, is_W0, 0, C.byref(cp))
        if not self.po:
            raise OfftError(f"offt_3d_init: {_err()}")

    # --- what the reference driver reads off the plan (run-fft.c:49-57, 320-321, 477-478) ---
    @property
    def comm(self) -> OfftComm:
        return self.po.contents.comm.contents

    @property
    def params(self) -> list:
        return list(self.po.contents.params.contents.v)

    @property
    def t(self) -> list:
        return list(self.po.contents.t)

    @property
    def rank(self) -> int:
        return self.po.contents.rank

    def box(self) -> dict:
        c = self.comm
        return {k: tuple(getattr(c, k)) for k in "istart isize istride ostart osize ostride".split()}

    @property
    def alloc_elems(self) -> int:
        return int(lib.offtb_plan_alloc_elems(self.po))

    @property
    def precision(self) -> int:
        return lib.offtb_plan_precision(self.po)

    def set_stream(self, stream_handle: int):
        _check(lib.offtb_plan_set_stream(self.po, stream_handle), "offtb_plan_set_stream")

    def set_async(self, on: bool):
        _check(lib.offtb_plan_set_async(self.po, int(on)), "offtb_plan_set_async")

    def set_stage_timing(self, on: bool):
        _check(lib.offtb_plan_set_stage_timing(self.po, int(on)), "offtb_plan_set_stage_timing")

    def execute(self, array):
        ptr = _ptr(array)
        lib.offtb_clear_error()
        lib.offt_3d_execute(self.po, ptr, ptr, 0)   # void, like the reference: a failure leaves a message
        if _err():
            raise OfftError(f"offt_3d_execute: {_err()}")

    def execute_inverse(self, array):
        ptr = _ptr(array)
        _check(lib.offt_3d_execute_inverse(self.po, ptr, ptr), "offt_3d_execute_inverse")

    def tune(self, array, max_loop=20, verbose=0) -> int:
        ptr = _ptr(array)
        n = lib.offtb_tune(self.po, ptr, ptr, max_loop, verbose)
        if n < 0:
            raise OfftError(f"offtb_tune: {_err()}")
        return n

    def tune_ex(self, max_loop=20, strategy=3, search_p1=False, verbose=0) -> int:
        """the search of tune.cu: strategy 0/1 Nelder-Mead from the reference's initial simplex, 2 random, 3 coordinate
        descent; search_p1 also moves the decomposition (read .comm / .alloc_elems again afterwards)"""
        n = lib.offtb_tune_ex(self.po, None, None, max_loop, verbose, strategy, int(search_p1))
        if n < 0:
            raise OfftError(f"offtb_tune_ex: {_err()}")
        return n

    def tune_harmony(self, max_loop=20, strategy=0, search_p1=False, verbose=0) -> int:
        """the same loop with the reference's Active Harmony server proposing the points where its back end was built
        (offt_b200/ah/_root), the built-in sources otherwise - what `run-fft -l` goes through"""
        n = lib.offtb_tune_harmony(self.po, max_loop, verbose, strategy, int(search_p1))
        if n < 0:
            raise OfftError(f"offtb_tune_harmony: {_err()}")
        return n

    @property
    def last_ms(self) -> float:
        return float(lib.offtb_plan_last_ms(self.po))

    @property
    def last_launches(self) -> int:
        return int(lib.offtb_plan_last_launches(self.po))

    def stage_ms(self) -> dict:
        buf = (C.c_double * 8)()
        lib.offtb_plan_stage_ms(self.po, buf, 8)
        return dict(zip(("k1_fftz", "k2_ffty", "k3_ffty", "k4_fftx", "exchange1", "exchange2", "h2d", "d2h"), buf))

    def exchange_block_elems(self, phase: int, myT: int) -> int:
        return int(lib.offtb_exchange_block_elems(self.po, phase, myT))

    def fin(self):
        if self.po:
            lib.offt_3d_fin(self.po)
            self.po = None

    def __del__(self):
        try:
            self.fin()
        except Exception:
            pass


def execute_group(plans, arrays, inverse=False):
    n = len(plans)
    pp = (C.POINTER(OfftPlan) * n)(*[p.po for p in plans])
    aa = (C.c_void_p * n)(*[_ptr(a) for a in arrays])
    _check(lib.offtb_execute_group(pp, aa, n, int(inverse)), "offtb_execute_group")


def fft_rows(data, n, stride, dist, howmany, sign=-1, bits=64, repeat=1, stream=0) -> float:
    ms = lib.offtb_fft_rows(_ptr(data), n, stride, dist, howmany, sign, bits, repeat, stream)
    if ms < 0:
        raise OfftError(f"offtb_fft_rows: {_err()}")
    return ms


def fft_launch_raw(src, dst, n, nbatch, im, om, *, bits=64, sign=-1, c_log=-1, load_cfast=0, store_cfast=0,
                   ry=(-1, 0, 0, 10), repeat=1, stream=0) -> float:
    im9 = (C.c_longlong * 9)(*im)
    om9 = (C.c_longlong * 9)(*om)
    ms = lib.offtb_fft_launch_raw(_ptr(src), _ptr(dst), n, bits, sign, nbatch, im9, om9, c_log, load_cfast, store_cfast,
                                  ry[0], ry[1], ry[2], ry[3], repeat, stream)
    if ms < 0:
        raise OfftError(f"offtb_fft_launch_raw: {_err()}")
    return ms


# ---- host logic, callable without a GPU ---------------------------------------------------------
def params_default(Nx, Ny, Nz, p, is_W0=0, is_notest=0) -> list:
    v = (C.c_int * PARAM_COUNT)()
    lib.offtb_params_default(Nx, Ny, Nz, p, is_W0, is_notest, v)
    return list(v)


def params_range(Nx, Ny, Nz, p) -> list:
    stride = 128
    lists = (C.c_int * (PARAM_COUNT * stride))()
    sizes = (C.c_int * PARAM_COUNT)()
    lib.offtb_params_range(Nx, Ny, Nz, p, lists, stride, sizes)
    return [list(lists[i * stride: i * stride + sizes[i]]) for i in range(PARAM_COUNT)]


def is_infeasible_point(Nx, Ny, Nz, p, v) -> tuple:
    vv = (C.c_int * PARAM_COUNT)(*v)
    bad = C.c_int(-1)
    r = lib.offtb_is_infeasible_point(Nx, Ny, Nz, p, vv, C.byref(bad))
    return int(r), int(bad.value)


def params_adjust(Nx, Ny, Nz, p, is_oned, v) -> list:
    vv = (C.c_int * PARAM_COUNT)(*v)
    lib.offtb_params_adjust(Nx, Ny, Nz, p, is_oned, vv)
    return list(vv)


def comm_box(Nx, Ny, Nz, p, p1, rank, S=0, is_equalxy=0, is_r2c=0) -> dict:
    c = OfftComm()
    _check(lib.offtb_comm_fill_r2c(C.byref(c), Nx, Ny, Nz, p, p1, rank, S, is_equalxy, is_r2c), "offtb_comm_fill")
    d = {k: getattr(c, k) for k in "p1 p2 M1 M2 M3 M4 F1 F2 F3 F4 m1 m2 m3 m4 b1 b2 b3 b4".split()}
    d.update({k: tuple(getattr(c, k)) for k in "istart isize istride ostart osize ostride".split()})
    return d


def alloc_elems(Nx, Ny, Nz, p, p1, is_r2c=0) -> int:
    return int(lib.offtb_alloc_elems_r2c(Nx, Ny, Nz, p, p1, is_r2c))


def check_supported(Nx, Ny, Nz, p, p1) -> tuple:
    rc = lib.offtb_check_supported(Nx, Ny, Nz, p, p1)
    return rc, (_err() if rc else "")
