"""Per-stage device times (po->t[]) of one single-GPU plan: python tools/plan_stages.py Nx Ny Nz [bits] [S]
OFFTB_LIB=<path> selects a variant build."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import offt_b200 as ob  # noqa: E402

N = tuple(int(v) for v in sys.argv[1:4])
bits = int(sys.argv[4]) if len(sys.argv) > 4 else 64
S = int(sys.argv[5]) if len(sys.argv) > 5 else 1
ob.world_fin(); ob.world_init_local(1, 0)
ob.set_default_precision(bits)
plan = ob.Plan(*N, is_notest=1, custom={ob.P.P1: 1, ob.P.S: S})
a = torch.zeros(plan.alloc_elems, dtype=torch.complex128 if bits == 64 else torch.complex64, device="cuda")
best = None
for _ in range(5):
    plan.execute(a)
    t = plan.t
    if best is None or t[0] < best[0]:
        best = t
names = {0: "ALL", 7: "FFTz", 8: "FFTy1", 9: "FFTy2", 10: "FFTx"}
gb = 2 * (16 if bits == 64 else 8) * N[0] * N[1] * N[2] / 1e9
print(f"lib={os.environ.get('OFFTB_LIB', 'default')} N={N} bits={bits} S={S}: " +
      "  ".join(f"{names[k]} {best[k] * 1e3:.3f} ms" + (f" ({gb / best[k] / 1e3:.2f} TB/s)" if k and best[k] > 0 else "") for k in names))
plan.fin()
