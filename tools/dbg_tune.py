import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.getcwd() + "/tests")
import offt_b200 as ob
P = ob.P
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
idt = torch.zeros(128, dtype=torch.uint8, device=dev)
if rank == 0: idt.copy_(torch.frombuffer(bytearray(ob.get_unique_id()), dtype=torch.uint8))
dist.broadcast(idt, 0)
ob.world_init(rank, world, local, idt.cpu().numpy().tobytes())
N = (128, 64, 128)
for S, T2, W2, timing in [(0, 64, 2, False), (0, 64, 2, True), (1, 64, 2, False), (1, 64, 2, True), (1, 32, 2, False), (0, 128, 0, True)]:
    plan = ob.Plan(*N, is_oned=1, is_notest=1, custom={P.P1: world, P.S: S, P.T2: T2, P.W2: W2})
    plan.set_stage_timing(timing)
    arr = torch.zeros(plan.alloc_elems, dtype=torch.complex128, device=dev)
    ok = True
    for i in range(3):
        try:
            plan.execute(arr)
        except Exception as e:
            print(f"rank {rank}: S={S} T2={T2} W2={W2} timing={timing} execute {i} FAILED {str(e)[:90]}", flush=True); ok = False; break
    if ok: print(f"rank {rank}: S={S} T2={T2} W2={W2} timing={timing} ok {plan.last_ms:.3f} ms", flush=True)
    plan.fin()
