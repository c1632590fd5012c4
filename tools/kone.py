"""launch one pass kind a few times (for ncu): python tools/kone.py <n> <z|y|x|xt> <c_log> [bits] [reps]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import offt_b200 as ob  # noqa: E402

n, mode, c_log = int(sys.argv[1]), sys.argv[2], int(sys.argv[3])
bits = int(sys.argv[4]) if len(sys.argv) > 4 else 64
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
ob.world_fin(); ob.world_init_local(1, 0)
a = torch.view_as_complex(torch.rand((n, n, n, 2), device="cuda", dtype=torch.float64 if bits == 64 else torch.float32)).contiguous()
b = torch.empty_like(a)
N2 = n * n
maps = {"z": ([0, 0, 0, 1, n, n, n, N2, 0], None, 0, 0, a), "y": ([0, 0, 0, n, n, 1, n, N2, 0], None, 1, 1, a),
        "x": ([0, 0, 0, N2, n, 1, n, n, 0], None, 1, 1, a), "xt": ([0, 0, 0, N2, n, 1, n, n, 0], [0, 0, 0, 1, n, N2, n, n, 0], 1, 0, b)}
im, om, lc, sc, dst = maps[mode]
om = om or im
ms = ob.fft_launch_raw(a, dst, n, N2, im, om, bits=bits, c_log=c_log, load_cfast=lc, store_cfast=sc, repeat=reps)
esz = 16 if bits == 64 else 8
print(f"{mode} n={n} c_log={c_log} bits={bits}: {ms:.3f} ms  {2 * esz * n ** 3 / ms / 1e6:.1f} GB/s")
