"""Time whole single-GPU plans that run on the any-length kernel (fft_generic.cu): grids with odd factors, the same
kernel forced onto a power-of-two grid, and a real-to-complex plan.  python tools/generic_bench.py"""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import offt_b200 as ob  # noqa: E402

P = ob.P
ob.world_fin(); ob.world_init_local(1, 0)


def run(N, r2c=0, force=False, label=""):
    ob.set_force_generic(force)
    plan = ob.Plan(*N, is_notest=1, is_r2c=r2c, custom={P.P1: 1, P.S: 1})
    plan.set_stage_timing(True)
    a = torch.view_as_complex(torch.rand((plan.alloc_elems, 2), device="cuda", dtype=torch.float64)).contiguous()
    for _ in range(3):
        plan.execute(a)
    n = N[0] * N[1] * N[2]
    ms = plan.last_ms
    st = {k: round(v, 3) for k, v in plan.stage_ms().items() if v > 0}
    print(f"{label or N}: {ms:.3f} ms  {5 * n * math.log2(n) / ms / 1e6:.0f} GFLOP/s  {6 * 16 * n / ms / 1e6:.0f} GB/s algorithmic (3 passes)  stages {st}", flush=True)
    plan.fin()
    ob.set_force_generic(False)


run((512, 512, 512), label="512^3 fast kernels")
run((512, 512, 512), force=True, label="512^3 forced onto the generic kernel")
run((384, 384, 384), label="384^3 (2^7*3)")
run((480, 480, 480), label="480^3 (2^5*3*5)")
run((500, 500, 500), label="500^3 (2^2*5^3)")
run((343, 343, 343), label="343^3 (7^3)")
run((509, 509, 509), label="509^3 (prime)")
run((512, 512, 512), r2c=1, label="512^3 real-to-complex")
