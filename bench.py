#!/usr/bin/env python
"""bench.py - the headline benchmark of BASELINE.json: forward 3-D complex FFT, GFLOP/s = 5*N*log2(N)/t.

    python bench.py --gpus 1 --steps K --warmup W            512^3 complex128 on one B200 (configs[1])
    torchrun ... bench.py --gpus N --steps K --warmup W      1024^3 complex128, slab, N = 2/4/8 (configs[2])
    python bench.py --impl reference ...                     the reference's own CPU path (oracle/_ref) on the host cores

One "step" is one forward transform (offt_3d_execute) of a seeded random grid.  The own arm prints
ONE JSON line (rank 0) with
  value      GFLOP/s, whole job, arrays resident in HBM when the timed region starts (device events,
             first kernel -> last kernel on the plan's compute stream, max over ranks per step);
  e2e        the same metric through the C API with HOST arrays: pinned host -> device copy, transform,
             device -> host copy, all inside the timed region (what run-fft.c's rep loop pays);
  roofline   the dominant kernel's algorithmic bytes (2 * 16 B per point per pass) over its own average
             duration (CUDA events around that launch on the launching stream) against the measured HBM peak;
  cpu_baseline  the unmodified reference (oracle/_ref/ref_dump: offt-compute.c over the stand-in MPI/FFTW)
             on the host cores, bounded sample, N=1 only.
  parity     (N >= 2) the gate run before the timed region through the same world: 256^3 slab and pencil plans,
             forward + backward, gathered on rank 0 and compared with numpy.fft.fftn through ostart/osize/ostride;
             and, at every N, 64 random (kx, ky) output pencils of the benchmark's own full-size result recomputed
             independently (DFT by definition over x and y as a float64 GEMM on the device, numpy FFT along z).
             Above 1e-12 the line carries "invalid" and the exit code is non-zero.
Nothing here falls back to the CPU: without the CUDA library the own arm fails.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

NVLINK_MEASURED_GBS = 770.0   # measured peer-copy bandwidth per direction on this pool (B200_PROFILING.md)
NVLINK_NOMINAL_GBS = 900.0    # NVLink 5 per direction, the figure BASELINE.json's north_star quotes
PARITY_TOL = 1e-12


def flops(N):
    n = N[0] * N[1] * N[2]
    return 5.0 * n * math.log2(n)


def workload_for(args):
    if args.grid:
        N = tuple(int(v) for v in args.grid.split("x"))
        if len(N) == 1:
            N = N * 3
        return N
    return (512, 512, 512) if args.gpus == 1 else (1024, 1024, 1024)


# This program's stdout carries exactly ONE JSON line.  Libraries underneath print there too (the reference's
# parameter lines from print_params, offt-compute.c:3416, 3469; NCCL's version banner), so file descriptor 1 is
# pointed at stderr for the whole run and the line goes out through the saved descriptor.
_REAL_STDOUT = None


def quiet_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: str):
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        print(line, flush=True)
    else:
        os.write(_REAL_STDOUT, (line + "\n").encode())


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 50 ms while the timed region runs"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []
        self.first = 0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        self.first = len(self.lines)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.lines[self.first:] or self.lines[-1:]:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------ reference arm
def host_ranks():
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        pass
    p = 1
    while p * 2 <= min(n, 64):
        p *= 2
    return p, n


def host_mem_available_gib():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) / 1048576.0
    except OSError:
        pass
    return 0.0


def run_reference_cpu(N, reps, p):
    """the unmodified reference pipeline (oracle/_ref/ref_dump) on p forked host ranks, slab P1 = p with -o
    (run-fft.c -o -d p): returns the per-rep seconds"""
    exe = ROOT / "oracle" / "_ref" / "ref_dump"
    if not exe.exists():
        return run_port_cpu(N, reps)
    p1 = p
    while N[0] % p1 or N[1] % p1:
        p1 //= 2
    cmd = [str(exe), str(N[0]), str(N[1]), str(N[2]), "1", "-", "1", "0", str(reps), f"0={p1}"]
    res = subprocess.run(cmd, env=dict(os.environ, OFFT_SHIM_NP=str(p1)), capture_output=True, text=True, timeout=3000)
    if res.returncode != 0:
        raise RuntimeError(f"ref_dump failed: {res.stdout[-500:]} {res.stderr[-500:]}")
    t = [float(line.split()[-1]) for line in res.stdout.splitlines() if line.startswith("ref_dump t_rep")]
    return t, p1


def run_port_cpu(N, reps):
    """fallback where the compiled reference is absent (it can only be built next to /root/reference): the C
    restatement of its pipeline (oracle/offt_oracle.c, built on demand with gcc), one thread, grid capped at 256^3.
    Returns per-rep seconds and 0 ranks (the callers label the run as a port)."""
    from oracle import oracle as O
    while N[0] * N[1] * N[2] > 256 ** 3:
        N = tuple(max(v // 2, 2) for v in N)
    orc = O.Oracle()
    grid = O.grid_values(1, *N)
    v = orc.resolve_params(*N, 1, {O.P1: 1})
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        orc.execute(grid, 1, v, 0, 0)
        t.append(time.perf_counter() - t0)
    run_port_cpu.grid = N
    return t, 0


def scipy_baseline(N, reps=3):
    """CPU Baseline B of BASELINE.md section 3: the best CPU FFT available in the image, scipy.fft.fftn (pocketfft)
    with every host core, on the same grid (bounded to 512^3)"""
    try:
        import numpy as np
        import scipy.fft
        _, cores = host_ranks()
        S = N
        while S[0] * S[1] * S[2] > 512 ** 3:
            S = tuple(max(v // 2, 2) for v in S)
        rng = np.random.default_rng(1)
        x = rng.uniform(-1, 1, S) + 1j * rng.uniform(-1, 1, S)
        t = []
        for _ in range(reps):
            t0 = time.perf_counter()
            scipy.fft.fftn(x, workers=cores)
            t.append(time.perf_counter() - t0)
        best = sum(t[1:]) / len(t[1:])
        return {"value": round(flops(S) / best / 1e9, 3), "unit": "GFLOP/s", "cores": cores, "kind": "scipy.fft.fftn(workers=cores)",
                "sample": f"{S[0]}x{S[1]}x{S[2]} complex128 forward, {reps} transforms (first untimed), out of place",
                "ms_per_step": round(best * 1e3, 2)}
    except Exception as e:   # noqa: BLE001 - a missing scipy must not hide the GPU numbers
        return {"value": None, "unit": "GFLOP/s", "cores": 0, "kind": "scipy.fft.fftn", "sample": f"unavailable: {e}"}


def cpu_baseline_of(N, reps=3):
    """the `cpu_baseline` object of the own arm: a bounded sample of the workload on the host cores (first transform untimed)"""
    p, cores = host_ranks()
    S = N
    while S[0] * S[1] * S[2] > 512 ** 3:
        S = tuple(max(v // 2, 2) for v in S)
    try:
        t, p_used = run_reference_cpu(S, reps, p)
        best = sum(t[1:]) / len(t[1:])
        if p_used == 0:   # no compiled reference on this box: the C port, one thread
            S = run_port_cpu.grid
            return {"value": round(flops(S) / best / 1e9, 3), "unit": "GFLOP/s", "cores": 1, "kind": "port",
                    "sample": f"{S[0]}x{S[1]}x{S[2]} complex128 forward, {reps} transforms (first untimed), oracle/offt_oracle.c on one thread "
                              "(oracle/_ref/ref_dump, the compiled reference, is absent on this box)",
                    "ms_per_step": round(best * 1e3, 2), "host_cores_visible": cores}
        return {"value": round(flops(S) / best / 1e9, 3), "unit": "GFLOP/s", "cores": p_used, "kind": "reference",
                "sample": f"{S[0]}x{S[1]}x{S[2]} complex128 forward, {reps} transforms (first untimed), slab {p_used}x1, "
                          f"{p_used} shim-MPI ranks of the unmodified reference (oracle/_ref/ref_dump)",
                "ms_per_step": round(best * 1e3, 2), "host_cores_visible": cores,
                "note": "reference pipeline over the stand-in MPI/FFT of oracle/shim (real MPI+FFTW cannot be installed here); "
                        "cpu_baseline_b is the best CPU FFT library in the image on the same grid"}
    except Exception as e:   # the checker being absent must not hide the GPU numbers
        return {"value": None, "unit": "GFLOP/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    N = workload_for(args)
    p, cores = host_ranks()
    # The workload itself where the host can hold it (1024^3: 16 GiB of grid + the reference's exchange buffers, about
    # 3-5 s per transform on 32 ranks); otherwise a bounded sample, and then config.workload names what actually ran.
    S = N
    need_gib = 2.6 * 16 * S[0] * S[1] * S[2] / 2 ** 30
    while S[0] * S[1] * S[2] > 512 ** 3 and (need_gib > 0.8 * host_mem_available_gib() or args.ref_sample):
        S = tuple(max(v // 2, 2) for v in S)
        need_gib = 2.6 * 16 * S[0] * S[1] * S[2] / 2 ** 30
    t, p_used = run_reference_cpu(S, args.warmup + args.steps, p)
    t = t[args.warmup:]
    ms = 1e3 * sum(t) / len(t)
    kind = "reference"
    if p_used == 0:      # the compiled reference is absent here: C port of its pipeline, one thread
        S, p_used, kind = run_port_cpu.grid, 1, "port"
    val = flops(S) / (ms * 1e-3) / 1e9
    sample = f"{S[0]}x{S[1]}x{S[2]} complex128 forward, " + (f"slab {p_used}x1 (-o -d {p_used}), {p_used} shim-MPI ranks; " if kind == "reference" else "oracle port on one thread; ") + \
             f"{'the whole workload' if S == N else 'a bounded sample of the ' + 'x'.join(map(str, N)) + ' workload'}"
    cfg = config_of(args, S)
    cfg["decomposition"] = f"host CPU: slab {p_used}x1 on {p_used} forked shim-MPI ranks" if kind == "reference" else "host CPU: one thread"
    if S != N:
        cfg["note"] = f"the own arm's workload is {N[0]}x{N[1]}x{N[2]}; this host could only run the grid named in `workload` (rates compare, times do not)"
    line = {"impl": "reference", "metric": "3D FFT GFLOP/s (5*N*log2(N)/t), forward, complex128", "value": round(val, 3),
            "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg, "gpu_launches": 0,
            "cpu_baseline": {"value": round(val, 3), "unit": "GFLOP/s", "cores": p_used, "kind": kind, "sample": sample,
                             "host_cores_visible": cores,
                             "note": "reference pipeline (offt-compute.c, unmodified) over the stand-in MPI and FFT of oracle/shim, not FFTW"
                                     if kind == "reference" else "C restatement of the reference pipeline (oracle/offt_oracle.c); the compiled reference is absent on this box"},
            "e2e": {"value": round(val, 3), "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(json.dumps(line))
    return 0


def config_of(args, N):
    n = args.gpus
    if n == 1:
        par = "single GPU, three local passes, no exchange"
    else:
        par = f"slab {n}x1 (is_oned, P1={n}), tiled all-to-all overlapped with the passes"
    return {"workload": f"{N[0]}x{N[1]}x{N[2]} complex128 forward 3-D FFT, in place ({'configs[1]' if N == (512,) * 3 else 'configs[2]' if N == (1024,) * 3 else 'custom grid'})",
            "decomposition": par,
            "l2": "arrays (>= 2 GiB per GPU) are larger than the 126 MB L2 and are re-filled from a pristine copy between steps"}


def traffic_of(N, kernel, args):
    """dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/traffic.json), or null"""
    if args.traffic is not None:
        return args.traffic
    try:
        t = json.loads((ROOT / "profiles" / "traffic.json").read_text())["bytes_per_launch"]
        return t.get(f"{N[0]}x{N[1]}x{N[2]}:{kernel}")
    except (OSError, ValueError, KeyError):
        return None


# ------------------------------------------------------------------------------------ own arm
class Ctx:
    """what every measurement below needs: the torch / distributed handles of this rank"""

    def __init__(self, torch, dist, ob, dev, rank, world):
        self.torch, self.dist, self.ob, self.dev, self.rank, self.world = torch, dist, ob, dev, rank, world

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.cpu().tolist()

    def sum_over_ranks(self, t):
        if self.world > 1:
            self.dist.all_reduce(t)
        return t


def prefer_numa_node_of_gpu(torch, local):
    """The host-array (e2e) path moves 2 x alloc bytes per GPU per step over PCIe: with all ranks' pinned buffers on one
    socket the other socket's GPUs cross the inter-socket link and one socket's memory controllers carry everything
    (round 1: 14.7 GB/s per GPU per direction at 8 GPUs against 52 at 1).  What `numactl --preferred` would do for the
    caller: prefer the NUMA node the GPU hangs off for this process's allocations.  Returns the node or None."""
    try:
        import ctypes
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        dev = torch.cuda.get_device_properties(local).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        mask = ctypes.c_ulong(1 << node)
        MPOL_PREFERRED, SYS_set_mempolicy = 1, 238   # x86_64
        rc = ctypes.CDLL(None, use_errno=True).syscall(SYS_set_mempolicy, MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(64))
        return node if rc == 0 else None
    except Exception:   # noqa: BLE001 - placement is an optimisation, never a reason to fail
        return None


def rel_l2(a, b):
    import numpy as np
    a = np.asarray(a).astype(np.complex128).ravel()
    b = np.asarray(b).astype(np.complex128).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def sampled_pencil_check(cx, plan, pristine, result, N, nsamp=64, seed=7):
    """Full-size parity without a full-size oracle: `nsamp` random output pencils X[kx, ky, :] of the result the
    benchmark itself produced, recomputed independently of the library - the sums over x and y by the DFT's
    definition (one float64 complex GEMM per rank over its input box, all-reduced), numpy's FFT along z - and
    compared in relative L2.  Reads the input through istart/isize/istride and the output through
    ostart/osize/ostride like the reference driver does (run-fft.c:49-57, 477-478)."""
    import numpy as np
    from offt_b200 import layout
    torch = cx.torch
    box = plan.box()
    Nx, Ny, Nz = N
    rng = np.random.default_rng(seed)
    kx = torch.from_numpy(rng.integers(0, Nx, nsamp)).to(cx.dev)
    ky = torch.from_numpy(rng.integers(0, Ny, nsamp)).to(cx.dev)
    cdt = pristine.dtype
    rdt = torch.float64
    xin = layout.input_view(box, pristine)                     # [sx, sy, Nz] strided view of this rank's input box
    sx, sy, sz = (int(v) for v in box["isize"])
    x0, y0 = int(box["istart"][0]), int(box["istart"][1])
    xs = torch.arange(x0, x0 + sx, device=cx.dev)
    ys = torch.arange(y0, y0 + sy, device=cx.dev)
    # exp(-2 pi i k n / N) with the phase reduced mod N in integers first
    ex = torch.polar(torch.ones(nsamp, sx, dtype=rdt, device=cx.dev), -2 * math.pi * ((kx[:, None] * xs[None, :]) % Nx).to(rdt) / Nx)
    ey = torch.polar(torch.ones(nsamp, sy, dtype=rdt, device=cx.dev), -2 * math.pi * ((ky[:, None] * ys[None, :]) % Ny).to(rdt) / Ny)
    part = torch.zeros(nsamp, Nz, dtype=torch.complex128, device=cx.dev)
    step = max(1, (1 << 24) // max(sy * sz, 1))                  # about 256 MiB of input per GEMM
    for a in range(0, sx, step):
        b = min(sx, a + step)
        w = (ex[:, a:b, None] * ey[:, None, :]).reshape(nsamp, (b - a) * sy)
        blk = xin[a:b].to(torch.complex128).reshape((b - a) * sy, sz)
        part[:, :sz] += w @ blk
        del w, blk
    part = cx.sum_over_ranks(torch.view_as_real(part))
    want = np.fft.fft(torch.view_as_complex(part).cpu().numpy(), axis=1)
    xout = layout.output_view(box, result)                    # [Nx, m4, m3]
    oy0, oz0 = int(box["ostart"][1]), int(box["ostart"][2])
    m4, m3 = int(box["osize"][1]), int(box["osize"][2])
    got = torch.zeros(nsamp, Nz, dtype=torch.complex128, device=cx.dev)
    mine = ((ky >= oy0) & (ky < oy0 + m4)).cpu().tolist()
    for s in range(nsamp):
        if mine[s]:
            got[s, oz0:oz0 + m3] = xout[int(kx[s]), int(ky[s]) - oy0, :].to(torch.complex128)
    got = torch.view_as_complex(cx.sum_over_ranks(torch.view_as_real(got))).cpu().numpy()
    del cdt
    return rel_l2(got, want)


def gate_case(cx, N, oned, custom, grid, want, bits=64):
    """one small plan through this world, forward and backward, every rank's array gathered and compared on rank 0
    with numpy (forward: through ostart/osize/ostride; backward: through istart/isize/istride, divided by N)"""
    import numpy as np
    from offt_b200 import layout
    torch, dist, ob = cx.torch, cx.dist, cx.ob
    P = ob.P
    plan = ob.Plan(*N, is_oned=oned, is_notest=1, custom=custom)
    box, alloc, params = plan.box(), plan.alloc_elems, plan.params
    arr = torch.from_numpy(layout.scatter_input(box, grid, alloc)).to(cx.dev)
    plan.execute(arr)
    fwd = arr.clone()
    plan.execute_inverse(arr)
    plan.fin()
    fw = [torch.empty_like(fwd) for _ in range(cx.world)]
    bw = [torch.empty_like(arr) for _ in range(cx.world)]
    dist.all_gather(fw, fwd)
    dist.all_gather(bw, arr)
    out = None
    if cx.rank == 0:
        boxes = [ob.comm_box(*N, cx.world, params[P.P1], r, params[P.S], 0) for r in range(cx.world)]
        e_f = rel_l2(layout.gather_output(boxes, [t.cpu().numpy() for t in fw], N), want)
        e_b = rel_l2(layout.gather_input(boxes, [t.cpu().numpy() for t in bw], N) / float(np.prod(N)), grid)
        out = {"grid": list(N), "process_grid": [params[P.P1], cx.world // params[P.P1]], "S": params[P.S],
               "forward_vs_numpy": e_f, "round_trip": e_b}
    del fw, bw, fwd, arr
    torch.cuda.empty_cache()
    return out


def parity_gate(cx):
    """the multi-process exchange path pinned where the driver can see it (N >= 2): 256^3 through the slab the
    benchmark uses, the other slab, and (4+ GPUs) pencils - forward and backward against numpy"""
    import numpy as np
    P = cx.ob.P
    N = (256, 256, 256)
    rng = np.random.default_rng(2024)
    grid = rng.uniform(-1, 1, N) + 1j * rng.uniform(-1, 1, N)
    want = np.fft.fftn(grid) if cx.rank == 0 else None
    w = cx.world
    cases = [(1, {P.P1: w, P.S: 0, P.T2: 32, P.W2: 3}), (1, {P.P1: 1, P.S: 1})]
    if w >= 4:
        cases += [(0, {P.P1: 2, P.S: 0}), (0, {P.P1: w // 2, P.S: 1, P.T1: 24, P.T2: 40})]
    res = [gate_case(cx, N, oned, custom, grid, want) for oned, custom in cases]
    return res


def measure_config(cx, N, *, bits, oned, custom, steps, tune=0, label=""):
    """one more BASELINE configuration through the same world (extra_configs): time, Parseval, round trip"""
    torch, ob = cx.torch, cx.ob
    P = ob.P
    ob.set_default_precision(bits)
    try:
        plan = ob.Plan(*N, is_oned=oned, is_notest=1, custom=custom)
        plan.set_stage_timing(False)
        alloc = plan.alloc_elems
        rdt = torch.float64 if bits == 64 else torch.float32
        g = torch.Generator(device=cx.dev); g.manual_seed(99 + cx.rank)
        x0 = torch.view_as_complex(torch.rand((alloc, 2), generator=g, device=cx.dev, dtype=rdt) * 2 - 1)
        w = torch.empty_like(x0)
        out = {"workload": label, "grid": list(N), "bits": bits}
        if tune > 0:
            # the default point first, then the library's search over T, W, Ry, S and the decomposition P1 (every trial
            # rebuilds the layout, offt-tuning.c:929-948, and runs on an internal zeroed array), then the arrays again
            out["default_params"] = {ob.PARAM_NAMES[i]: plan.params[i] for i in (P.P1, P.T1, P.W1, P.T2, P.W2, P.Ry, P.S)}
            for _ in range(3):
                w.copy_(x0); cx.barrier(); plan.execute(w)
            out["default_ms"] = round(cx.max_over_ranks([plan.last_ms])[0], 4)
            del x0, w
            torch.cuda.empty_cache()
            out["tune_trials"] = plan.tune_ex(tune, strategy=3, search_p1=True)
            plan.set_stage_timing(False)
            alloc = plan.alloc_elems
            g = torch.Generator(device=cx.dev); g.manual_seed(99 + cx.rank)
            x0 = torch.view_as_complex(torch.rand((alloc, 2), generator=g, device=cx.dev, dtype=rdt) * 2 - 1)
            w = torch.empty_like(x0)
        times = []
        for i in range(steps + 2):
            w.copy_(x0); cx.barrier()
            plan.execute(w)
            if i >= 2:
                times.append(cx.max_over_ranks([plan.last_ms])[0])
        fl = flops(N)
        out.update({"params": {ob.PARAM_NAMES[i]: plan.params[i] for i in (P.P1, P.T1, P.W1, P.T2, P.W2, P.Ry, P.S)},
                    "ms_per_step": round(sum(times) / len(times), 4), "ms_min": round(min(times), 4),
                    "GFLOPs": round(fl / (sum(times) / len(times)) / 1e6, 1), "steps": steps})
        e = torch.stack([x0.real.double().square().sum() + x0.imag.double().square().sum(),
                         w.real.double().square().sum() + w.imag.double().square().sum()])
        e = cx.sum_over_ranks(e)
        out["parseval_rel_err"] = abs(float(e[1]) / (float(e[0]) * N[0] * N[1] * N[2]) - 1.0)
        if bits == 64:
            out["sampled_pencils_rel_l2"] = sampled_pencil_check(cx, plan, x0, w, N, nsamp=16)
        plan.execute_inverse(w)
        w /= float(N[0] * N[1] * N[2])
        d = torch.stack([(w - x0).abs().double().square().sum(), x0.abs().double().square().sum()])
        d = cx.sum_over_ranks(d)
        out["round_trip_rel_l2"] = float(torch.sqrt(d[0] / d[1]))
        del x0, w
        plan.fin()
        torch.cuda.empty_cache()
        return out
    except Exception as e:   # noqa: BLE001 - an extra configuration must not take the headline line down
        return {"workload": label, "error": str(e)[-300:]}
    finally:
        ob.set_default_precision(64)


def own_arm(args):
    import torch
    import torch.distributed as dist
    import offt_b200 as ob   # loads offt_b200/lib/libofft_b200.so; raises if it is not built

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product has no CPU path")
    quiet_stdout()
    os.environ.setdefault("OFFTB_FLAG_TIMEOUT_S", "60")   # a lost peer fails the run with a message instead of hanging the box
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    N = workload_for(args)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(ob.get_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        ob.world_init(rank, world, local, bytes(idt.cpu().numpy().tobytes()))
    else:
        ob.world_init(0, 1, local, None)
    cx = Ctx(torch, dist, ob, dev, rank, world)
    barrier = cx.barrier

    # ---- parity gate of the multi-process path, before anything is timed
    gate = parity_gate(cx) if world > 1 and not args.no_gate else None

    P = ob.P
    # tunables: the reference leaves S, T2, W2 to its tuner; these are the values the sweep in DESIGN.md picked
    S = args.S if args.S >= 0 else (1 if world == 1 else 0)
    custom = {P.P1: world, P.S: S}
    T2 = args.T2 if args.T2 > 0 else (64 if world > 1 and N[2] >= 256 else 0)
    if T2 > 0:
        custom[P.T2] = T2
    W2 = args.W2 if args.W2 >= 0 else (3 if world > 1 and N[2] >= 256 else -1)   # profiles/r01_cfg3_sweep_1024_8gpu.json
    if W2 >= 0:
        custom[P.W2] = W2
    plan = ob.Plan(*N, is_oned=1 if world > 1 else 0, is_notest=1, custom=custom)
    alloc = plan.alloc_elems
    nbytes = alloc * 16
    # seeded synthetic grid, generated on the device per rank (local box of a global random grid)
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    pristine = torch.view_as_complex((torch.rand((alloc, 2), generator=g, device=dev, dtype=torch.float64) * 2 - 1))
    work = torch.empty_like(pristine)

    def one_step():
        work.copy_(pristine)          # untimed re-fill (also evicts the previous result from L2)
        barrier()
        plan.execute(work)            # synchronous; device time first->last kernel in plan.last_ms
        return plan.last_ms, plan.last_launches

    plan.set_stage_timing(False)  # the timed steps carry no per-launch events; the per-kernel pass below turns them on
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()            # nvidia-smi needs a few hundred ms to come up: start it before the warm-up
    for _ in range(max(args.warmup, 3)):
        one_step()
    barrier()
    clocks.mark()                 # only samples taken from here on (timed steps + per-kernel timing) are reported
    wall0 = time.perf_counter()
    times, launches = [], 0
    for _ in range(args.steps):
        ms, l = one_step()
        times.append(ms); launches += l
    barrier()
    wall = time.perf_counter() - wall0
    step_ms = cx.max_over_ranks(times)
    if world > 1:
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt)
        launches = int(lt.item())
    ms_per_step = sum(step_ms) / len(step_ms)
    value = flops(N) / (ms_per_step * 1e-3) / 1e9

    # ---- size-independent property check of the last result (Parseval): |X|^2 = N |x|^2
    et = torch.stack([pristine.real.square().sum() + pristine.imag.square().sum(), work.real.square().sum() + work.imag.square().sum()])
    e_in, e_out = cx.sum_over_ranks(et).tolist()
    parseval = abs(e_out / (e_in * N[0] * N[1] * N[2]) - 1.0)
    # ---- and the full-size result itself against an independent computation of 64 of its pencils
    pencils = sampled_pencil_check(cx, plan, pristine, work, N)

    # ---- per-kernel durations (CUDA events around every launch on its stream) -> roofline
    plan.set_stage_timing(True)
    acc, nrep = {}, 5
    for _ in range(nrep):
        one_step()
        for k, v in plan.stage_ms().items():
            acc[k] = acc.get(k, 0.0) + v / nrep
    plan.set_stage_timing(False)
    clock_rec = clocks.stop() if rank == 0 else None
    # launches per step of each fused kernel: 1 at N=1; one per tile for the tiled phase
    params = plan.params
    c = plan.comm
    tiles2 = -(-c.m3 // params[P.T2]) if world > 1 else 1
    pass_bytes = 2 * 16 * (N[0] * N[1] * N[2]) / world          # one read + one write of the local data
    kern = {"k1_fftz": 1, "k2_ffty": 1 if world == 1 else 0, "k3_ffty": tiles2 if world > 1 else 0, "k4_fftx": tiles2 if world > 1 else 1}
    chained = world > 1 and os.environ.get("OFFTB_PDL", "1") != "0" and os.environ.get("OFFTB_OVERLAP", "1") != "0" and os.environ.get("OFFTB_EXCHANGE", "") != "nccl"
    passes = {}
    for k, nl in kern.items():
        if nl and acc.get(k, 0) > 0:
            passes[k] = {"ms_per_step": round(acc[k], 4), "launches_per_step": nl, "GBps": round(pass_bytes / (acc[k] * 1e-3) / 1e9, 1)}
            if nl > 1 and chained:
                passes[k]["timed_as"] = "span of the stream's dependent-launch chain (first launch start -> last launch end, flag waits included)"
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak, peak_src = float(json.loads(peaks_file.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    # The dominant HBM-bound kernel.  One GPU: the slowest of the three passes.  Several GPUs: the z pass K1 - the
    # chains K3 (writer) and K4 (reader) run side by side paced by NVLink, their spans include flag waits and are
    # reported under `exchange` and `passes`, not against the HBM roofline.
    hbm_passes = {k: v for k, v in passes.items() if world == 1 or k == "k1_fftz"} or passes
    dom = max(hbm_passes, key=lambda k: hbm_passes[k]["ms_per_step"])
    achieved = passes[dom]["GBps"]
    roofline = {"bound": "hbm", "kernel": f"fft_kernel<double> as {dom}", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": int(pass_bytes / passes[dom]["launches_per_step"]),
                "avg_launch_ms": round(passes[dom]["ms_per_step"] / passes[dom]["launches_per_step"], 5),
                "traffic": traffic_of(N, dom, args), "passes": passes}
    exchange = None
    if world > 1:
        xb = (world - 1) / world * 16 * (N[0] * N[1] * N[2]) / world      # bytes each GPU sends (and receives)
        xms = acc.get("exchange2", 0.0)
        fused = xms == 0.0     # fused exchange: the NVLink stores are part of the K3 launches, there is no separate copy
        if fused:
            xms = acc.get("k3_ffty", 0.0)
        rate = xb / (xms * 1e-3) / 1e9 if xms > 0 else None
        exchange = {"bytes_out_per_gpu": int(xb), "device_ms": round(xms, 4),
                    "how": ("stores of the K3 (FFTy + pack) launches into the peers' slots over NVLink; " +
                            ("device_ms = span of the writer chain" if chained else "device_ms = sum over the tiles' launches")) if fused else "grouped ncclSend/ncclRecv",
                    "GBps_per_direction": round(rate, 1) if rate else None,
                    "frac_of_900": round(rate / NVLINK_NOMINAL_GBS, 4) if rate else None,
                    "frac_of_770": round(rate / NVLINK_MEASURED_GBS, 4) if rate else None,
                    "peaks": "900 = NVLink 5 nominal per direction (north_star); 770 = measured peer copy per direction (B200_PROFILING.md)"}

    # ---- e2e: host arrays through the C API (H2D + transform + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        numa_node = prefer_numa_node_of_gpu(torch, local) if world > 1 else None
        host = torch.empty(alloc, dtype=torch.complex128, pin_memory=True)
        keep = torch.empty(alloc, dtype=torch.complex128, pin_memory=True)
        keep.copy_(pristine)
        e_steps = max(3, min(args.steps, 5))
        e_times = []
        for i in range(2 + e_steps):
            host.copy_(keep)
            barrier()
            t0 = time.perf_counter()
            plan.execute(host)        # cudaMemcpyAsync H2D, kernels, cudaMemcpyAsync D2H, stream sync
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if i >= 2:
                e_times.append(dt * 1e3)
        e_ms = sum(cx.max_over_ranks(e_times)) / len(e_times)
        e2e = {"value": round(flops(N) / (e_ms * 1e-3) / 1e9, 2), "unit": "GFLOP/s", "ms_per_step": round(e_ms, 3), "steps": e_steps,
               "h2d_bytes_per_step": int(nbytes) * world, "d2h_bytes_per_step": int(nbytes) * world,
               "timing": "host wall clock around offt_3d_execute(host pointer) + synchronize, max over ranks",
               "host_numa_node_rank0": numa_node}
        del host, keep

    cpu_baseline = cpu_baseline_of(N) if rank == 0 and world == 1 and not args.no_cpu else None
    cpu_baseline_b = scipy_baseline(N) if rank == 0 and world == 1 and not args.no_cpu else None

    plan.fin()
    del pristine, work
    torch.cuda.empty_cache()

    # ---- more of BASELINE.json's configurations through the same world
    extra = {}
    extra_configs = []
    if not args.no_extra:
        if world == 1 and N == (512, 512, 512):
            # configs[2]'s grid on ONE GPU: the missing point of a true strong-scaling curve (1 -> 2/4/8 GPUs on 1024^3)
            extra["strong_scaling_n1_1024"] = measure_config(cx, (1024, 1024, 1024), bits=64, oned=0, custom={P.P1: 1, P.S: 1}, steps=5,
                                                             label="1024^3 complex128 forward on one GPU (configs[2]'s grid without the exchange)")
        if args.extra_test and world > 1:
            # the same code path as the 8-GPU extras on grids that fit any world (validation runs of this script)
            p1x = 2 if world >= 4 else world
            extra_configs.append(measure_config(cx, (256, 256, 256), bits=64, oned=0 if world >= 4 else 1, custom={P.P1: p1x, P.S: 0}, steps=3,
                                                label=f"extra-test: 256^3 complex128, P1={p1x}"))
            extra_configs.append(measure_config(cx, (512, 256, 128), bits=32, oned=1, custom={P.P1: world, P.S: 0}, steps=3, tune=10,
                                                label="extra-test: 512x256x128 complex64, tunables searched"))
        if world == 8 and N == (1024, 1024, 1024):
            extra_configs.append(measure_config(cx, (2048, 2048, 2048), bits=64, oned=0, custom={P.P1: 2, P.S: 0}, steps=3,
                                                label="configs[3]: 2048^3 complex128 forward, pencil 2x4, 8 GPUs"))
            extra_configs.append(measure_config(cx, (2048, 1024, 512), bits=32, oned=1, custom={P.P1: 8, P.S: 0}, steps=5, tune=24,
                                                label="configs[4]: 2048x1024x512 complex64 forward, 8 GPUs, tile / window / P1 searched (offtb_tune_ex, coordinate descent incl. the decomposition)"))

    ob.world_fin()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    parity = {"tolerance": PARITY_TOL, "sampled_pencils": {"count": 64, "grid": list(N), "rel_l2": pencils,
                                                         "how": "64 random (kx, ky) pencils of the timed workload's own result vs the DFT by definition over x, y (float64 GEMM) + numpy FFT along z"},
              "cases": gate}
    worst = [pencils] + [v for cse in (gate or []) for v in (cse["forward_vs_numpy"], cse["round_trip"])]
    parity["rel_l2"] = max(worst)
    line = {"metric": "3D FFT GFLOP/s (5*N*log2(N)/t), forward, complex128", "value": round(value, 2), "unit": "GFLOP/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4),
            "ms_min": round(min(step_ms), 4), "ms_median": round(statistics.median(step_ms), 4),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(config_of(args, N), tunables={ob.PARAM_NAMES[i]: params[i] for i in (P.P1, P.T1, P.W1, P.T2, P.W2, P.Ry, P.V, P.S)}),
            "timing": "CUDA events on the plan's compute stream around each step (first kernel -> last kernel, exchanges included), max over ranks per step; barrier + synchronize on both sides of every step",
            "wall_s_timed_region": round(wall, 3), "gpu_launches": launches,
            "parseval_rel_err": parseval, "parity": parity, "clocks": clock_rec, "roofline": roofline, "exchange": exchange,
            "e2e": e2e, "cpu_baseline": cpu_baseline, "cpu_baseline_b": cpu_baseline_b}
    if extra:
        line["extra"] = extra
    if extra_configs:
        line["extra_configs"] = extra_configs
    bad = parseval > 1e-9 or not (parity["rel_l2"] <= PARITY_TOL)
    if bad:
        line["invalid"] = f"parity check failed (Parseval {parseval:.3e}, worst rel-L2 {parity['rel_l2']:.3e} > {PARITY_TOL})"
    emit(json.dumps(line))
    return 1 if bad else 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--grid", default="", help="NxMxL or N (default: 512^3 at 1 GPU, 1024^3 otherwise)")
    ap.add_argument("--S", type=int, default=-1, help="tunable _S_: 1 = x-y-z output (the reference's STRIDE mode); 0 = z-y-x (its default); -1: 1 on one GPU, 0 on several")
    ap.add_argument("--T2", type=int, default=0, help="tile thickness of the exchange phase (0: reference default)")
    ap.add_argument("--W2", type=int, default=-1, help="overlap window (-1: reference default)")
    ap.add_argument("--traffic", type=float, default=None, help="dram bytes per launch of the dominant kernel from an ncu capture (profiles/)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-gate", action="store_true", help="skip the multi-process parity gate (experiments only)")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra configurations (1024^3 on one GPU; configs[3], configs[4] at 8 GPUs)")
    ap.add_argument("--extra-test", action="store_true", help="run the extra-configuration code path on small grids (validation of this script)")
    ap.add_argument("--ref-sample", action="store_true", help="reference arm: force the bounded 512^3 sample even if the host could run the workload")
    args = ap.parse_args()
    sys.exit(reference_arm(args) if args.impl == "reference" else own_arm(args))


if __name__ == "__main__":
    main()
