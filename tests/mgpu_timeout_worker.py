"""One rank of the flag time-out check (torchrun, 2+ GPUs): rank 1 stays away from one execute, so the other ranks'
kernels wait for flags that never come.  With OFFTB_FLAG_TIMEOUT_S set the waiting CTAs must leave, offt_3d_execute must
fail with a message (no trap, no hang), and after the plan is destroyed the same processes must run a new plan correctly."""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import offt_b200 as ob  # noqa: E402
from offt_b200 import layout  # noqa: E402

P = ob.P


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
    N = (64, 64, 64)
    plan = ob.Plan(*N, is_oned=1, is_notest=1, custom={P.P1: world, P.T2: 32})   # two tiles: two waits at most
    arr = torch.zeros(plan.alloc_elems, dtype=torch.complex128, device=dev)
    msg = ""
    t0 = time.time()
    if rank != 1:
        try:
            plan.execute(arr)
        except ob.OfftError as e:
            msg = str(e)
    else:
        time.sleep(float(os.environ.get("OFFTB_FLAG_TIMEOUT_S", "3")) + 2.0)
    waited = time.time() - t0
    plan.fin()            # collective: the failed plan is destroyed by everybody
    # the context survived: a fresh plan transforms correctly
    rng = np.random.default_rng(5)
    grid = rng.uniform(-1, 1, N) + 1j * rng.uniform(-1, 1, N)
    plan = ob.Plan(*N, is_oned=1, is_notest=1, custom={P.P1: world})
    box = plan.box()
    a = torch.from_numpy(layout.scatter_input(box, grid, plan.alloc_elems)).to(dev)
    plan.execute(a)
    outs = [torch.empty_like(a) for _ in range(world)]
    dist.all_gather(outs, a)
    plan.fin()
    ok = True
    if rank == 0:
        boxes = [ob.comm_box(*N, world, world, r, 0, 0) for r in range(world)]
        got = layout.gather_output(boxes, [t.cpu().numpy() for t in outs], N)
        err = float(np.linalg.norm(got - np.fft.fftn(grid)) / np.linalg.norm(np.fft.fftn(grid)))
        ok = "timed out" in msg and waited < 60 and err < 1e-12
        print(f"{'ok  ' if ok else 'FAIL'} time-out: rank 0 got '{msg[:70]}...' after {waited:.1f} s; next plan vs numpy {err:.2e}", flush=True)
    ob.world_fin()
    ft = torch.tensor([0 if ok else 1], device=dev)
    dist.broadcast(ft, 0)
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU TIMEOUT", "PASSED" if int(ft.item()) == 0 else "FAILED", flush=True)
    sys.exit(int(ft.item()))


if __name__ == "__main__":
    main()
