"""TEST INFRASTRUCTURE - not product code.

Python face of the oracle for the distributed 3-D FFT path of rchyena/offt
(`offt_3d_execute`, offt-compute.c:3864).  It offers

* `grid_values`      - the seeded synthetic grid every checker and the product's
                       bench generate identically,
* `Oracle`           - ctypes binding of `oracle/_ref/liboracle.so`, the C
                       restatement of the reference pipeline (offt_oracle.c),
* `run_reference`    - runs the UNMODIFIED reference (`oracle/_ref/ref_dump`,
                       built from /root/reference by oracle/Makefile) on the
                       host cores and returns every rank's descriptor + output,
* `gather_output` / `scatter_input` - index through istart/isize/istride and
                       ostart/osize/ostride exactly as run-fft.c:46-61, 452-503.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import
this module; the product path (offt_b200/) never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import tempfile
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"
PARAM_COUNT = 24
# parameter indices, offt.h:74-98
P1, T1, W1, PX1, PY1, FZ, FP1, UX1, UZ1, FU1, FY1, RY, T2, W2, PZ2, PX2, FY2, FP2, UZ2, UY2, FU2, FX, V, S = range(24)


def _splitmix(seed: int, index: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
             + index.astype(np.uint64) * np.uint64(0xD1B54A32D192ED03)
             + np.uint64(0x632BE59BD9B4E019))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 / 9007199254740992.0) - 1.0


def grid_values(seed: int, Nx: int, Ny: int, Nz: int, x0: int = 0, x1: int | None = None) -> np.ndarray:
    """complex128 [x1-x0, Ny, Nz] slab of the seeded global grid: element with global
    linear index g = (x*Ny + y)*Nz + z has re = u(seed, 2g), im = u(seed, 2g+1), u in [-1, 1)."""
    x1 = Nx if x1 is None else x1
    g = np.arange(x0 * Ny * Nz, x1 * Ny * Nz, dtype=np.uint64)
    re = _splitmix(seed, g * np.uint64(2))
    im = _splitmix(seed, g * np.uint64(2) + np.uint64(1))
    return (re + 1j * im).reshape(x1 - x0, Ny, Nz)


def ramp_values(Nx: int, Ny: int, Nz: int) -> np.ndarray:
    """run-fft.c:46-61 input pattern: in[x,y,z] = z + 10*y + 100*x (global coordinates), imag 0."""
    x = np.arange(Nx, dtype=np.float64)[:, None, None]
    y = np.arange(Ny, dtype=np.float64)[None, :, None]
    z = np.arange(Nz, dtype=np.float64)[None, None, :]
    return (z + 10.0 * y + 100.0 * x).astype(np.complex128)


@dataclass
class RankBox:
    """What `struct _offt_comm` (offt.h:102-142) tells the caller about one rank."""
    p: int
    rank: int
    N: tuple
    p1: int
    p2: int
    istart: tuple
    isize: tuple
    istride: tuple
    ostart: tuple
    osize: tuple
    ostride: tuple
    alloc: int                      # complex elements of the in-place array (run-fft.c:294-304)
    params: list = field(default_factory=list)
    data: np.ndarray | None = None  # complex128 [alloc]


def scatter_input(box: RankBox, grid: np.ndarray) -> np.ndarray:
    """rank-local in-place array holding `grid`'s input box (run-fft.c:49-57)."""
    a = np.zeros(box.alloc, dtype=grid.dtype)
    sx, sy, sz = box.isize
    ix = np.arange(sx)[:, None, None] * box.istride[0]
    iy = np.arange(sy)[None, :, None] * box.istride[1]
    iz = np.arange(sz)[None, None, :] * box.istride[2]
    sub = grid[box.istart[0]:box.istart[0] + sx, box.istart[1]:box.istart[1] + sy, box.istart[2]:box.istart[2] + sz]
    a[(ix + iy + iz).ravel()] = sub.ravel()
    return a


def gather_output(boxes: list, dtype=np.complex128) -> np.ndarray:
    """global [Nx, Ny, Nz] spectrum assembled through ostart/osize/ostride (run-fft.c:477-478)."""
    Nx, Ny, Nz = boxes[0].N
    out = np.full((Nx, Ny, Nz), np.nan + 0j, dtype=dtype)
    for b in boxes:
        sx, sy, sz = b.osize
        if sx * sy * sz == 0:
            continue
        ix = np.arange(sx)[:, None, None] * b.ostride[0]
        iy = np.arange(sy)[None, :, None] * b.ostride[1]
        iz = np.arange(sz)[None, None, :] * b.ostride[2]
        out[b.ostart[0]:b.ostart[0] + sx, b.ostart[1]:b.ostart[1] + sy, b.ostart[2]:b.ostart[2] + sz] = \
            b.data[(ix + iy + iz).ravel()].reshape(sx, sy, sz)
    return out


def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    """|a-b|_2 / |b|_2, the parity measure BASELINE.json's north_star names."""
    a = np.asarray(a).astype(np.complex128).ravel()
    b = np.asarray(b).astype(np.complex128).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def have_reference() -> bool:
    return (REF_DIR / "ref_dump").exists()


def scatter_input_r2c(box: RankBox, real_grid: np.ndarray) -> np.ndarray:
    """in-place r2c input of one rank as a complex128 view of its double array: real (x, y, z) of the box sits at
    double index z + 2*(istride1*y + istride0*x) (run-fft.c:53-55)"""
    a = np.zeros(2 * box.alloc, dtype=np.float64)
    sx, sy, sz = box.isize
    ix = np.arange(sx)[:, None, None] * (2 * box.istride[0])
    iy = np.arange(sy)[None, :, None] * (2 * box.istride[1])
    iz = np.arange(sz)[None, None, :]
    sub = real_grid[box.istart[0]:box.istart[0] + sx, box.istart[1]:box.istart[1] + sy, box.istart[2]:box.istart[2] + sz]
    a[(ix + iy + iz).ravel()] = sub.ravel()
    return a.view(np.complex128)


def gather_input_r2c(boxes: list, arrays: list) -> np.ndarray:
    """global real grid re-assembled from every rank's r2c-layout array (what the backward transform returns)"""
    Nx, Ny, Nz = boxes[0].N
    out = np.full((Nx, Ny, Nz), np.nan)
    for b, a in zip(boxes, arrays):
        d = np.ascontiguousarray(a).view(np.float64)
        sx, sy, sz = b.isize
        ix = np.arange(sx)[:, None, None] * (2 * b.istride[0])
        iy = np.arange(sy)[None, :, None] * (2 * b.istride[1])
        iz = np.arange(sz)[None, None, :]
        out[b.istart[0]:b.istart[0] + sx, b.istart[1]:b.istart[1] + sy, b.istart[2]:b.istart[2] + sz] = d[(ix + iy + iz).ravel()].reshape(sx, sy, sz)
    return out


def gather_output_r2c(boxes: list, dtype=np.complex128) -> np.ndarray:
    """global [Nx, Ny, Nz/2+1] half spectrum through ostart/osize/ostride"""
    Nx, Ny, Nz = boxes[0].N
    out = np.full((Nx, Ny, Nz // 2 + 1), np.nan + 0j, dtype=dtype)
    for b in boxes:
        sx, sy, sz = b.osize
        if sx * sy * sz == 0:
            continue
        ix = np.arange(sx)[:, None, None] * b.ostride[0]
        iy = np.arange(sy)[None, :, None] * b.ostride[1]
        iz = np.arange(sz)[None, None, :] * b.ostride[2]
        out[b.ostart[0]:b.ostart[0] + sx, b.ostart[1]:b.ostart[1] + sy, b.ostart[2]:b.ostart[2] + sz] = \
            b.data[(ix + iy + iz).ravel()].reshape(sx, sy, sz)
    return out


def run_reference(Nx, Ny, Nz, p, seed, is_oned=0, is_equalxy=0, reps=1, params=None, keep_data=True, is_r2c=0):
    """Run the unmodified reference on `p` forked host ranks; returns (boxes, t_min seconds)."""
    params = dict(params or {})
    if P1 not in params:
        raise ValueError("params must set P1 (index 0)")
    with tempfile.TemporaryDirectory() as td:
        prefix = os.path.join(td, "dump") if keep_data else "-"
        cmd = [str(REF_DIR / "ref_dump"), str(Nx), str(Ny), str(Nz), str(seed), prefix,
               str(int(is_oned)), str(int(is_equalxy)), str(int(reps))]
        cmd += [f"{k}={v}" for k, v in sorted(params.items())]
        if is_r2c:
            cmd.append("r2c=1")
        env = dict(os.environ, OFFT_SHIM_NP=str(p))
        res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=3600)
        if res.returncode != 0:
            raise RuntimeError(f"ref_dump failed ({res.returncode}): {res.stdout[-2000:]} {res.stderr[-2000:]}")
        tmin = None
        for line in res.stdout.splitlines():
            if line.startswith("ref_dump t_min"):
                tmin = float(line.split()[-1])
        boxes = []
        if keep_data:
            for r in range(p):
                with open(f"{prefix}.rank{r}.bin", "rb") as f:
                    hdr = np.frombuffer(f.read(32 * 8), dtype=np.int64)
                    pv = np.frombuffer(f.read(PARAM_COUNT * 4), dtype=np.int32)
                    data = np.frombuffer(f.read(), dtype=np.float64)
                alloc = int(hdr[25])
                boxes.append(RankBox(
                    p=int(hdr[0]), rank=int(hdr[1]), N=tuple(int(v) for v in hdr[2:5]), p1=int(hdr[5]), p2=int(hdr[6]),
                    istart=tuple(int(v) for v in hdr[7:10]), isize=tuple(int(v) for v in hdr[10:13]),
                    istride=tuple(int(v) for v in hdr[13:16]), ostart=tuple(int(v) for v in hdr[16:19]),
                    osize=tuple(int(v) for v in hdr[19:22]), ostride=tuple(int(v) for v in hdr[22:25]),
                    alloc=alloc, params=[int(v) for v in pv],
                    data=data.view(np.complex128).copy()))
    return boxes, tmin


# ----------------------------------------------------------------------------
# ctypes face of oracle/_ref/liboracle.so (offt_oracle.c)
# ----------------------------------------------------------------------------
MAX_GRID = 128
_OCOMM_FIELDS = ("p1 p2 rank_x rank_y M1 M2 M3 M4 F1 F2 F3 F4 m1 m2 m3 m4 b1 b2 b3 b4").split()


def build_oracle(force: bool = False) -> Path:
    """compile offt_oracle.c -> _ref/liboracle.so (gcc only; runs on the GPU box too)."""
    so = REF_DIR / "liboracle.so"
    src = HERE / "offt_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < max(src.stat().st_mtime, (HERE / "oracle_dft.h").stat().st_mtime):
        subprocess.run(["make", "-C", str(HERE), "oracle"], check=True, capture_output=True)
    return so


class Oracle:
    def __init__(self):
        self.lib = ctypes.CDLL(str(build_oracle()))
        L = self.lib
        L.oracle_alloc_elems.restype = ctypes.c_int64
        L.oracle_alloc_elems.argtypes = [ctypes.c_int] * 5
        L.oracle_comm_fill.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 8
        L.oracle_execute.restype = ctypes.c_int
        L.oracle_execute.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.oracle_dft_rows.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int]
        L.oracle_params_range.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_params_default.argtypes = [ctypes.c_int] * 6 + [ctypes.c_void_p]
        L.oracle_is_infeasible.restype = ctypes.c_int
        L.oracle_is_infeasible.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_params_adjust.argtypes = [ctypes.c_int] * 5 + [ctypes.c_void_p]

    def alloc_elems(self, Nx, Ny, Nz, p, p1) -> int:
        return int(self.lib.oracle_alloc_elems(Nx, Ny, Nz, p, p1))

    def comm(self, Nx, Ny, Nz, p, p1, rank, S=0, is_equalxy=0) -> dict:
        buf = (ctypes.c_int * 38)()
        self.lib.oracle_comm_fill(buf, Nx, Ny, Nz, p, p1, rank, S, is_equalxy)
        v = list(buf)
        d = dict(zip(_OCOMM_FIELDS, v[:20]))
        for i, k in enumerate(("istart", "isize", "istride", "ostart", "osize", "ostride")):
            d[k] = tuple(v[20 + 3 * i: 23 + 3 * i])
        return d

    def box(self, Nx, Ny, Nz, p, p1, rank, S=0, is_equalxy=0) -> RankBox:
        d = self.comm(Nx, Ny, Nz, p, p1, rank, S, is_equalxy)
        return RankBox(p=p, rank=rank, N=(Nx, Ny, Nz), p1=d["p1"], p2=d["p2"], istart=d["istart"], isize=d["isize"],
                       istride=d["istride"], ostart=d["ostart"], osize=d["osize"], ostride=d["ostride"],
                       alloc=self.alloc_elems(Nx, Ny, Nz, p, p1))

    def params_default(self, Nx, Ny, Nz, p, is_W0=0, is_notest=0) -> list:
        v = (ctypes.c_int * PARAM_COUNT)()
        self.lib.oracle_params_default(Nx, Ny, Nz, p, is_W0, is_notest, v)
        return list(v)

    def params_range(self, Nx, Ny, Nz, p) -> list:
        lists = (ctypes.c_int * (PARAM_COUNT * MAX_GRID))()
        sizes = (ctypes.c_int * PARAM_COUNT)()
        self.lib.oracle_params_range(Nx, Ny, Nz, p, lists, sizes)
        return [list(lists[i * MAX_GRID: i * MAX_GRID + sizes[i]]) for i in range(PARAM_COUNT)]

    def is_infeasible(self, Nx, Ny, Nz, p, v) -> tuple:
        vv = (ctypes.c_int * PARAM_COUNT)(*v)
        bad = ctypes.c_int(-1)
        r = self.lib.oracle_is_infeasible(Nx, Ny, Nz, p, vv, ctypes.byref(bad))
        return int(r), int(bad.value)

    def params_adjust(self, Nx, Ny, Nz, p, is_oned, v) -> list:
        vv = (ctypes.c_int * PARAM_COUNT)(*v)
        self.lib.oracle_params_adjust(Nx, Ny, Nz, p, is_oned, vv)
        return list(vv)

    def resolve_params(self, Nx, Ny, Nz, p, custom=None, is_W0=0, is_notest=1) -> list:
        """defaults, then the caller's non-negative overrides (offt-compute.c:3415-3417)."""
        v = self.params_default(Nx, Ny, Nz, p, is_W0, is_notest)
        for k, val in (custom or {}).items():
            if val >= 0:
                v[k] = val
        return v

    def execute(self, grid: np.ndarray, p: int, v: list, is_oned=0, is_equalxy=0) -> list:
        """forward 3-D FFT of the global `grid` through the restated pipeline on p simulated
        ranks; returns one RankBox per rank with `.data` = that rank's in-place array."""
        Nx, Ny, Nz = grid.shape
        boxes = [self.box(Nx, Ny, Nz, p, v[P1], r, v[S], is_equalxy) for r in range(p)]
        arrays = [np.ascontiguousarray(scatter_input(b, grid.astype(np.complex128))) for b in boxes]
        ptrs = (ctypes.c_void_p * p)(*[a.ctypes.data for a in arrays])
        vv = (ctypes.c_int * PARAM_COUNT)(*v)
        rc = self.lib.oracle_execute(Nx, Ny, Nz, p, vv, is_oned, is_equalxy, ptrs)
        if rc != 0:
            raise ValueError("oracle_execute rejected the arguments")
        for b, a in zip(boxes, arrays):
            b.data = a
            b.params = list(v)
        return boxes

    def dft_rows(self, data: np.ndarray, n: int, stride: int, dist: int, howmany: int, sign: int = -1) -> None:
        assert data.dtype == np.complex128 and data.flags.c_contiguous
        self.lib.oracle_dft_rows(data.ctypes.data, n, stride, dist, howmany, sign)
