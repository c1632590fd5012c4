"""Run one BASELINE.json configuration through the C API on the GPUs torchrun gives us, check the size-independent
properties (Parseval, forward->backward returns N*x) and print one JSON line with times and the per-stage breakdown.

  torchrun --nproc-per-node 8 tools/run_config.py --grid 2048 --p1 2 --steps 3                      # configs[3]: pencil 2x4
  torchrun --nproc-per-node 8 tools/run_config.py --grid 2048x1024x512 --bits 32 --oned 1 --sweep   # configs[4]: T/W sweep
"""
import argparse
import json
import math
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import offt_b200 as ob  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", default="1024")
    ap.add_argument("--p1", type=int, default=0, help="process grid p1 (0: all ranks = slab p x 1)")
    ap.add_argument("--oned", type=int, default=-1, help="is_oned (-1: 1 for slabs, 0 for pencils)")
    ap.add_argument("--bits", type=int, default=64)
    ap.add_argument("--S", type=int, default=0)
    ap.add_argument("--T1", type=int, default=0); ap.add_argument("--W1", type=int, default=-1)
    ap.add_argument("--T2", type=int, default=0); ap.add_argument("--W2", type=int, default=-1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--sweep", action="store_true", help="time every (T, W) of the reference's value grid for the active phase(s)")
    ap.add_argument("--tune", type=int, default=0, help="offtb_tune with this many trials (the library's fetch/measure/report loop)")
    a = ap.parse_args()
    N = tuple(int(v) for v in a.grid.split("x"))
    N = N * 3 if len(N) == 1 else N
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(ob.get_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        ob.world_init(rank, world, local, idt.cpu().numpy().tobytes())
    else:
        ob.world_init(0, 1, local, None)
    ob.set_default_precision(a.bits)
    P = ob.P
    p1 = a.p1 or world
    oned = a.oned if a.oned >= 0 else int(p1 in (1, world))
    rdt = torch.float64 if a.bits == 64 else torch.float32
    flop = 5.0 * N[0] * N[1] * N[2] * math.log2(N[0] * N[1] * N[2])

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def agree(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def measure(custom, steps, check):
        plan = ob.Plan(*N, is_oned=oned, is_notest=1, custom=custom)
        alloc = plan.alloc_elems
        g = torch.Generator(device=dev); g.manual_seed(99 + rank)
        x0 = torch.view_as_complex(torch.rand((alloc, 2), generator=g, device=dev, dtype=rdt) * 2 - 1)
        w = torch.empty_like(x0)
        if a.tune > 0 and check:
            w.copy_(x0)
            plan.tune(w, a.tune, verbose=int(rank == 0))   # the fetch / measure / report loop of ah_tuning (offt-tuning.c:879-1006)
        times = []
        for i in range(steps + 2):
            w.copy_(x0); sync()
            plan.execute(w)
            if i >= 2:
                times.append(agree(plan.last_ms))
        out = {"params": {ob.PARAM_NAMES[i]: plan.params[i] for i in (P.P1, P.T1, P.W1, P.T2, P.W2, P.Ry, P.S)},
               "ms_min": round(min(times), 4), "ms_mean": round(sum(times) / len(times), 4), "GFLOPs": round(flop / min(times) / 1e6, 1)}
        if check:
            plan.set_stage_timing(True)
            w.copy_(x0); sync(); plan.execute(w)
            out["stages_ms"] = {k: round(v, 4) for k, v in plan.stage_ms().items() if v > 0}
            plan.set_stage_timing(False)
            e = torch.stack([x0.real.double().square().sum() + x0.imag.double().square().sum(),
                             w.real.double().square().sum() + w.imag.double().square().sum()])
            if world > 1:
                dist.all_reduce(e)
            out["parseval_rel_err"] = abs(float(e[1]) / (float(e[0]) * N[0] * N[1] * N[2]) - 1.0)
            plan.execute_inverse(w)
            w /= float(N[0] * N[1] * N[2])
            d = torch.stack([(w - x0).abs().double().square().sum(), x0.abs().double().square().sum()])
            if world > 1:
                dist.all_reduce(d)
            out["round_trip_rel_l2"] = float(torch.sqrt(d[0] / d[1]))
        del x0, w
        plan.fin()
        torch.cuda.empty_cache()
        return out

    base = {P.P1: p1, P.S: a.S}
    for k, v in ((P.T1, a.T1), (P.T2, a.T2)):
        if v > 0:
            base[k] = v
    for k, v in ((P.W1, a.W1), (P.W2, a.W2)):
        if v >= 0:
            base[k] = v
    res = {"grid": N, "bits": a.bits, "gpus": world, "process_grid": [p1, world // p1], "is_oned": oned,
           "main": measure(base, a.steps, True)}
    if a.sweep:
        # the active phase's T and W on the reference's value grid (params_range_setup, offt-compute.c:2998-3093)
        grid = ob.params_range(*N, world)
        phases = [(P.T2, P.W2)] if (oned and p1 == world) else [(P.T1, P.W1)] if (oned and p1 == 1) else [(P.T1, P.W1), (P.T2, P.W2)]
        sweep = []
        for Tk, Wk in phases:
            for T in [t for t in grid[Tk] if t >= 4]:
                for W in (0, 1, 2, 3):
                    c = dict(base); c[Tk] = T; c[Wk] = W
                    try:
                        m = measure(c, 2, False)
                        sweep.append({"T": T, "W": W, "phase": 1 if Tk == P.T1 else 2, "ms_min": m["ms_min"]})
                    except ob.OfftError as e:
                        sweep.append({"T": T, "W": W, "error": str(e)[-80:]})
        res["sweep"] = sweep
        ok = [s for s in sweep if "ms_min" in s]
        res["best"] = min(ok, key=lambda s: s["ms_min"]) if ok else None
    ob.world_fin()
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
