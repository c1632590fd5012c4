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

NVLINK_GBS = 770.0   # measured peer-copy bandwidth per direction on this pool (B200_PROFILING.md); nominal 900


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
    import numpy as np
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
                "note": "reference pipeline over the stand-in MPI/FFT of oracle/shim (real MPI+FFTW cannot be installed here)"}
    except Exception as e:   # the checker being absent must not hide the GPU numbers
        return {"value": None, "unit": "GFLOP/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    N = workload_for(args)
    p, cores = host_ranks()
    # bounded sample: the reference pipeline costs ~1.5 s per 256^3 on one core; keep a step at a few seconds
    S = N
    while S[0] * S[1] * S[2] > 512 ** 3:
        S = tuple(max(v // 2, 2) for v in S)
    t, p_used = run_reference_cpu(S, args.warmup + args.steps, p)
    t = t[args.warmup:]
    ms = 1e3 * sum(t) / len(t)
    kind = "reference"
    if p_used == 0:      # the compiled reference is absent here: C port of its pipeline, one thread
        S, p_used, kind = run_port_cpu.grid, 1, "port"
    val = flops(S) / (ms * 1e-3) / 1e9
    sample = f"{S[0]}x{S[1]}x{S[2]} complex128 forward, " + (f"slab {p_used}x1 (-o -d {p_used}), {p_used} shim-MPI ranks; " if kind == "reference" else "oracle port on one thread; ") + \
             f"{'the whole workload' if S == N else 'a bounded sample of the ' + 'x'.join(map(str, N)) + ' workload'}"
    line = {"impl": "reference", "metric": "3D FFT GFLOP/s (5*N*log2(N)/t), forward, complex128", "value": round(val, 3),
            "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(args, N), "gpu_launches": 0,
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
    return {"workload": f"{N[0]}x{N[1]}x{N[2]} complex128 forward 3-D FFT, in place ({'configs[1]' if n == 1 and N == (512,) * 3 else 'configs[2]' if N == (1024,) * 3 else 'custom grid'})",
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
def own_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import offt_b200 as ob   # loads offt_b200/lib/libofft_b200.so; raises if it is not built

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product has no CPU path")
    quiet_stdout()
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

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        work.copy_(pristine)          # untimed re-fill (also evicts the previous result from L2)
        barrier()
        plan.execute(work)            # synchronous; device time first->last kernel in plan.last_ms
        return plan.last_ms, plan.last_launches

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
    tt = torch.tensor(times, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt)
        launches = int(lt.item())
    step_ms = tt.cpu().tolist()
    ms_per_step = sum(step_ms) / len(step_ms)
    value = flops(N) / (ms_per_step * 1e-3) / 1e9

    # ---- size-independent property check of the last result (Parseval): |X|^2 = N |x|^2
    e_in = float((pristine.real.square().sum() + pristine.imag.square().sum()).item())
    e_out = float((work.real.square().sum() + work.imag.square().sum()).item())
    if world > 1:
        et = torch.tensor([e_in, e_out], dtype=torch.float64, device=dev)
        dist.all_reduce(et)
        e_in, e_out = et.tolist()
    parseval = abs(e_out / (e_in * N[0] * N[1] * N[2]) - 1.0)

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
    passes = {}
    for k, nl in kern.items():
        if nl and acc.get(k, 0) > 0:
            passes[k] = {"ms_per_step": round(acc[k], 4), "launches_per_step": nl, "GBps": round(pass_bytes / (acc[k] * 1e-3) / 1e9, 1)}
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak, peak_src = float(json.loads(peaks_file.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    dom = max(passes, key=lambda k: passes[k]["ms_per_step"])
    achieved = passes[dom]["GBps"]
    roofline = {"bound": "hbm", "kernel": f"fft_kernel<double> as {dom}", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": int(pass_bytes / passes[dom]["launches_per_step"]),
                "avg_launch_ms": round(passes[dom]["ms_per_step"] / passes[dom]["launches_per_step"], 5),
                "traffic": traffic_of(N, dom, args), "passes": passes}
    if world > 1:
        xb = (world - 1) / world * 16 * (N[0] * N[1] * N[2]) / world      # bytes each GPU sends (and receives)
        xms = acc.get("exchange2", 0.0)
        fused = xms == 0.0     # fused exchange: the NVLink stores are part of the K3 launches, there is no separate copy
        if fused:
            xms = acc.get("k3_ffty", 0.0)
        roofline["exchange"] = {"bytes_out_per_gpu": int(xb), "device_ms_sum_of_tiles": round(xms, 4),
                                "how": "stores of the K3 (FFTy + pack) launches into the peers' slots" if fused else "grouped ncclSend/ncclRecv",
                                "GBps_per_direction": round(xb / (xms * 1e-3) / 1e9, 1) if xms > 0 else None,
                                "peak": NVLINK_GBS, "frac": round(xb / (xms * 1e-3) / 1e9 / NVLINK_GBS, 4) if xms > 0 else None,
                                "peak_source": "measured peer copy per direction (B200_PROFILING.md); nominal 900"}

    # ---- e2e: host arrays through the C API (H2D + transform + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
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
        et = torch.tensor(e_times, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(et, op=dist.ReduceOp.MAX)
        e_ms = float(et.mean().item())
        e2e = {"value": round(flops(N) / (e_ms * 1e-3) / 1e9, 2), "unit": "GFLOP/s", "ms_per_step": round(e_ms, 3), "steps": e_steps,
               "h2d_bytes_per_step": int(nbytes) * world, "d2h_bytes_per_step": int(nbytes) * world,
               "timing": "host wall clock around offt_3d_execute(host pointer) + synchronize, max over ranks"}
        del host, keep

    cpu_baseline = cpu_baseline_of(N) if rank == 0 and world == 1 and not args.no_cpu else None

    plan.fin()
    ob.world_fin()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    line = {"metric": "3D FFT GFLOP/s (5*N*log2(N)/t), forward, complex128", "value": round(value, 2), "unit": "GFLOP/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4),
            "ms_min": round(min(step_ms), 4), "ms_median": round(statistics.median(step_ms), 4),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(config_of(args, N), tunables={ob.PARAM_NAMES[i]: params[i] for i in (P.P1, P.T1, P.W1, P.T2, P.W2, P.Ry, P.V, P.S)}),
            "timing": "CUDA events on the plan's compute stream around each step (first kernel -> last kernel, exchanges included), max over ranks per step; barrier + synchronize on both sides of every step",
            "wall_s_timed_region": round(wall, 3), "gpu_launches": launches,
            "parseval_rel_err": parseval, "clocks": clock_rec, "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu_baseline}
    if parseval > 1e-9:
        line["invalid"] = f"Parseval check failed ({parseval:.3e})"
    emit(json.dumps(line))
    return 0 if parseval <= 1e-9 else 1


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
    args = ap.parse_args()
    sys.exit(reference_arm(args) if args.impl == "reference" else own_arm(args))


if __name__ == "__main__":
    main()
