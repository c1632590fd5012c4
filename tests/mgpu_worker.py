"""One rank of the multi-GPU parity run (launched by torchrun, one process per GPU, NCCL).

Every rank builds the plan through the C API, fills its input box of the seeded global grid,
runs offt_3d_execute on its own GPU, and rank 0 gathers all outputs and compares them through
ostart/osize/ostride with the oracle and numpy.  Then the backward transform must return N * input.
The exchange path is whatever the environment selects (tests/test_multi_gpu.py runs them all):
default = fused peer stores (TMA bulk stores, two streams, dependent-launch chains); OFFTB_BULK=0 /
OFFTB_PDL=0 / OFFTB_OVERLAP=0 switch those pieces off; OFFTB_EXCHANGE=nccl = grouped ncclSend/ncclRecv.
usage: torchrun ... tests/mgpu_worker.py
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import offt_b200 as ob  # noqa: E402
from oracle import oracle as O  # noqa: E402
from gpu_helpers import box_of, gather_input  # noqa: E402

P = ob.P


def cases(p):
    out = [((64, 32, 128), 1, {P.P1: p}, 64), ((64, 32, 128), 1, {P.P1: 1}, 64), ((64, 64, 64), 1, {P.P1: p, P.S: 1, P.T2: 8, P.W2: 3}, 64),
           ((32, 64, 64), 1, {P.P1: 1, P.S: 1, P.T1: 5, P.W1: 1}, 64), ((256, 256, 256), 1, {P.P1: p, P.S: 1}, 64),
           ((128, 64, 64), 1, {P.P1: p}, 32),
           # lengths with odd factors and uneven splits: the any-length kernel and the general block split between real ranks
           ((15, 10, 9), 1, {P.P1: p, P.V: 3, P.T2: 2}, 64), ((27, 20, 45), 1, {P.P1: 1, P.S: 1, P.T1: 4}, 64),
           # in place with unequal plane strides on both sides of phase 1 at every p (Ny = 5p+1: planes of 6p rows in the
           # caller's layout, 5p+1 between the phases) - the backward transform depends on the tile order, plan.cu run_phase
           ((27, 5 * p + 1, 45), 1, {P.P1: 1, P.S: 1, P.T1: 2, P.W1: 1}, 64)]
    if p >= 4:
        out += [((10, 9, 15), 0, {P.P1: 2, P.V: 3, P.T1: 2, P.T2: 3}, 64), ((64, 64, 128), 0, {P.P1: 2}, 64), ((64, 128, 64), 0, {P.P1: p // 2, P.S: 1, P.Ry: 3}, 64), ((64, 64, 64), 0, {P.P1: 2, P.T1: 3, P.T2: 5}, 32),
                ((12, 5 * (p // 2) + 1, 15), 0, {P.P1: 2, P.S: 1, P.T1: 2, P.T2: 4}, 64)]
    return out


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    idt = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(ob.get_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    ob.world_init(rank, world, local, idt.cpu().numpy().tobytes())
    orc = O.Oracle()
    failures = 0
    for N, oned, custom, bits in cases(world):
        ob.set_default_precision(bits)
        cdt = np.complex128 if bits == 64 else np.complex64
        tol = 1e-12 if bits == 64 else 1e-5
        grid = O.grid_values(11, *N)
        plan = ob.Plan(*N, is_oned=oned, is_notest=1, custom=custom)
        box = box_of(plan, N, world)
        arr = torch.from_numpy(np.ascontiguousarray(O.scatter_input(box, grid.astype(cdt)))).to(dev)
        plan.execute(arr)
        launches = plan.last_launches
        fwd = arr.cpu().numpy()
        plan.execute_inverse(arr)
        back = arr.cpu().numpy()
        plan.fin()
        gathered = [None] * world
        dist.all_gather_object(gathered, (box, fwd, back))
        if rank == 0:
            boxes = []
            for b, f, _ in gathered:
                b.data = f
                boxes.append(b)
            got = O.gather_output(boxes, cdt)
            want = O.gather_output(orc.execute(grid, world, orc.resolve_params(*N, world, custom), oned, 0))
            e1 = O.rel_l2(got, want)
            e2 = O.rel_l2(got, np.fft.fftn(grid))
            rt = gather_input(boxes, [g[2] for g in gathered], cdt) / np.prod(N)
            e3 = O.rel_l2(rt, grid)
            ok = e1 < tol and e2 < tol and e3 < tol and launches > 0 and not np.isnan(got).any()
            failures += 0 if ok else 1
            print(f"{'ok  ' if ok else 'FAIL'} N={N} p={world} oned={oned} custom={custom} bits={bits}: vs oracle {e1:.2e}  vs numpy {e2:.2e}  "
                  f"round trip {e3:.2e}  launches/rank {launches}", flush=True)
    # real-to-complex plans between real ranks: slabs take the half-length fast path for their local z pass, phase-1 z
    # passes (slab 1 x p, pencils) the any-length kernel; forward vs numpy.fft.rfftn, backward returns N * the real input
    ob.set_default_precision(64)
    r2c_cases = [((64, 32, 128), 1, {P.P1: world}), ((32, 64, 250), 1, {P.P1: 1, P.S: 1})]
    if world >= 4:
        r2c_cases.append(((64, 64, 96), 0, {P.P1: 2, P.T1: 8, P.T2: 5}))
    for N, oned, custom in r2c_cases:
        grid = O.grid_values(13, *N)
        plan = ob.Plan(*N, is_oned=oned, is_notest=1, is_r2c=1, custom=custom)
        box = box_of(plan, N, world)
        arr = torch.from_numpy(np.ascontiguousarray(O.scatter_input_r2c(box, np.ascontiguousarray(grid.real)))).to(dev)
        plan.execute(arr)
        fwd = arr.cpu().numpy()
        plan.execute_inverse(arr)
        back = arr.cpu().numpy()
        plan.fin()
        gathered = [None] * world
        dist.all_gather_object(gathered, (box, fwd, back))
        if rank == 0:
            boxes = []
            for b, f, _ in gathered:
                b.data = f
                boxes.append(b)
            e1 = O.rel_l2(O.gather_output_r2c(boxes), np.fft.rfftn(grid.real))
            e2 = O.rel_l2(O.gather_input_r2c(boxes, [g[2] for g in gathered]) / np.prod(N), grid.real)
            ok = e1 < 1e-12 and e2 < 1e-12
            failures += 0 if ok else 1
            print(f"{'ok  ' if ok else 'FAIL'} r2c N={N} p={world} oned={oned} custom={custom}: vs numpy rfftn {e1:.2e}  round trip {e2:.2e}", flush=True)
    # the tuning loop (ah_tuning's fetch -> measure -> report, offt-tuning.c:879-1006) on the device: the plan must
    # come back with a feasible point on the reference's grid, identical on every rank, and still transform correctly
    ob.set_default_precision(64)
    N, oned, custom = (128, 64, 128), 1, {P.P1: world}
    grid = O.grid_values(3, *N)
    plan = ob.Plan(*N, is_oned=oned, is_notest=1, custom=custom)
    box = box_of(plan, N, world)
    arr = torch.from_numpy(np.ascontiguousarray(O.scatter_input(box, grid))).to(dev)
    trials = plan.tune(arr, 8)
    tuned = plan.params
    arr.copy_(torch.from_numpy(np.ascontiguousarray(O.scatter_input(box, grid))))
    plan.execute(arr)
    fwd = arr.cpu().numpy()
    plan.fin()
    gathered = [None] * world
    dist.all_gather_object(gathered, (box, fwd, tuned))
    if rank == 0:
        boxes = []
        for b, f, _ in gathered:
            b.data = f
            boxes.append(b)
        same = all(g[2] == tuned for g in gathered)
        rng = ob.params_range(*N, world)
        on_grid = all(tuned[k] in rng[k] for k in (P.T1, P.W1, P.T2, P.W2))
        e = O.rel_l2(O.gather_output(boxes), np.fft.fftn(grid))
        ok = same and on_grid and e < 1e-12 and trials >= 1
        failures += 0 if ok else 1
        print(f"{'ok  ' if ok else 'FAIL'} tuning: {trials} trials -> T2 {tuned[P.T2]} W2 {tuned[P.W2]}, same on all ranks {same}, on grid {on_grid}, vs numpy {e:.2e}", flush=True)
    # the same loop searching the decomposition too, Nelder-Mead from the reference's initial simplex (ah_strategy 0): P1 may
    # change, so the layout is read only afterwards (run-fft.c:314 -> :269-304) and the transform must still be right
    N, oned = (64, 128, 64), 0
    grid = O.grid_values(4, *N)
    plan = ob.Plan(*N, is_oned=oned, is_notest=1, custom={P.P1: 1})
    trials = plan.tune_ex(10, strategy=0, search_p1=True)
    tuned = plan.params
    box = box_of(plan, N, world)
    arr = torch.from_numpy(np.ascontiguousarray(O.scatter_input(box, grid))).to(dev)
    plan.execute(arr)
    fwd = arr.cpu().numpy()
    plan.fin()
    gathered = [None] * world
    dist.all_gather_object(gathered, (box, fwd, tuned))
    if rank == 0:
        boxes = []
        for b, f, _ in gathered:
            b.data = f
            boxes.append(b)
        same = all(g[2] == tuned for g in gathered)
        e = O.rel_l2(O.gather_output(boxes), np.fft.fftn(grid))
        ok = same and e < 1e-12 and trials >= 1
        failures += 0 if ok else 1
        print(f"{'ok  ' if ok else 'FAIL'} tuning incl. P1 (Nelder-Mead): {trials} trials -> P1 {tuned[P.P1]} T1 {tuned[P.T1]} W1 {tuned[P.W1]} T2 {tuned[P.T2]} W2 {tuned[P.W2]} S {tuned[P.S]}, "
              f"same on all ranks {same}, vs numpy {e:.2e}", flush=True)
    # and through the reference's own Active Harmony server (hserver + patched nm.so, built unmodified by offt_b200/ah)
    # where that back end travelled with the repo; elsewhere the same call falls back to the built-in Nelder-Mead
    N, oned = (64, 64, 128), 0
    grid = O.grid_values(8, *N)
    plan = ob.Plan(*N, is_oned=oned, is_notest=1, custom={P.P1: world})
    ah_built = (ROOT / "offt_b200" / "ah" / "_root" / "lib" / "libofft_ah.so").exists()
    trials = plan.tune_harmony(8, strategy=0, search_p1=True, verbose=int(rank == 0))
    tuned = plan.params
    box = box_of(plan, N, world)
    arr = torch.from_numpy(np.ascontiguousarray(O.scatter_input(box, grid))).to(dev)
    plan.execute(arr)
    fwd = arr.cpu().numpy()
    plan.fin()
    gathered = [None] * world
    dist.all_gather_object(gathered, (box, fwd, tuned))
    if rank == 0:
        boxes = []
        for b, f, _ in gathered:
            b.data = f
            boxes.append(b)
        same = all(g[2] == tuned for g in gathered)
        e = O.rel_l2(O.gather_output(boxes), np.fft.fftn(grid))
        ok = same and e < 1e-12 and trials >= 1
        failures += 0 if ok else 1
        print(f"{'ok  ' if ok else 'FAIL'} tuning through Active Harmony ({'hserver + nm.so' if ah_built else 'back end not built: built-in fallback'}): {trials} trials -> "
              f"P1 {tuned[P.P1]} T1 {tuned[P.T1]} W1 {tuned[P.W1]} T2 {tuned[P.T2]} W2 {tuned[P.W2]} S {tuned[P.S]}, same on all ranks {same}, vs numpy {e:.2e}", flush=True)
    ob.world_fin()
    ft = torch.tensor([failures], device=dev)
    dist.broadcast(ft, 0)
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU PARITY", "PASSED" if failures == 0 else f"FAILED ({failures})", flush=True)
    sys.exit(int(ft.item()) != 0)


if __name__ == "__main__":
    main()
