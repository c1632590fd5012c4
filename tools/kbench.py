"""Per-pass timing of the local passes of an N^3 complex FFT on one GPU (device-resident, CUDA events
around `repeat` back-to-back launches).  Prints GB/s = 2*B*N^3 / t per pass.

    python tools/kbench.py [n] [bits] [--modes z,y,x,xt] [--clogs -1,0,1,2,3] [--copy] [--plan]
--copy times the same address maps with the kernel's move-only path (the Ry rule switched to "never transform"):
the ceiling of the access pattern itself.  OFFTB_LIB=<path> selects a variant build (tools/variants.sh)."""
import argparse
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import offt_b200 as ob  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("n", type=int, nargs="?", default=512)
    ap.add_argument("bits", type=int, nargs="?", default=64)
    ap.add_argument("--modes", default="z,y,x,xt")
    ap.add_argument("--clogs", default="-1,0,1,2,3")
    ap.add_argument("--copy", action="store_true")
    ap.add_argument("--plan", action="store_true")
    ap.add_argument("--rep", type=int, default=10)
    ap.add_argument("--xpad", default="", help="comma list of pads (elements) added to the x stride: modes xp (in place), xpl (padded load, dense store), xps (dense load, padded store)")
    a = ap.parse_args()
    n, bits, rep = a.n, a.bits, a.rep
    ob.world_fin(); ob.world_init_local(1, 0)
    rows = min(n * n, (1 << 27) // n)   # keep the array at 2 GiB (f64) for n > 512
    x = torch.view_as_complex(torch.rand((rows * n, 2), device="cuda", dtype=torch.float64 if bits == 64 else torch.float32)).contiguous()
    y = torch.empty_like(x)
    esz = 16 if bits == 64 else 8
    gb = 2 * esz * rows * n / 1e9
    nb = rows                      # batch = rows; the other two axes are (n, rows / n)
    r1 = rows // n
    # element (i, j, k) of an [n, r1... ] cube: z: k fastest
    maps = {
        "z": ([0, 0, 0, 1, n, n, r1, n * n, 0], None, 0, 0, x),
        "y": ([0, 0, 0, n, n, 1, r1, n * n, 0], None, 1, 1, x),
        "x": ([0, 0, 0, rows, n, 1, r1, n, 0], None, 1, 1, x),
        "xt": ([0, 0, 0, rows, n, 1, r1, n, 0], [0, 0, 0, 1, n, rows, r1, n * n, 0], 1, 0, y),
        # transform along an axis while swapping the two outer axes: loads at the row stride, stores at the plane stride ...
        "xs": ([0, 0, 0, n, n, 1, r1, n * n, 0], [0, 0, 0, rows, n, 1, r1, n, 0], 1, 1, y),
        # ... and the other way round
        "sx": ([0, 0, 0, rows, n, 1, r1, n, 0], [0, 0, 0, n, n, 1, r1, n * n, 0], 1, 1, y),
    }
    pads = [int(v) for v in a.xpad.split(",")] if a.xpad else []
    if pads:
        big = torch.empty(rows * n + n * max(pads), device="cuda", dtype=x.dtype)
        big.copy_(torch.cat([x, x[: n * max(pads)]]))
        for pd in pads:
            xs = rows + pd
            maps[f"xp{pd}"] = ([0, 0, 0, xs, n, 1, r1, n, 0], None, 1, 1, big, big)
            maps[f"xpl{pd}"] = ([0, 0, 0, xs, n, 1, r1, n, 0], [0, 0, 0, rows, n, 1, r1, n, 0], 1, 1, y, big)
            maps[f"xps{pd}"] = ([0, 0, 0, rows, n, 1, r1, n, 0], [0, 0, 0, xs, n, 1, r1, n, 0], 1, 1, big, x)
        a.modes = ",".join(m for m in a.modes.split(",") if m) + "," + ",".join(k for k in maps if k.startswith("xp"))
        a.modes = a.modes.strip(",")
    print(f"lib={os.environ.get('OFFTB_LIB', 'default')} N={n} rows={rows} bits={bits}: {gb:.2f} GB per pass (algorithmic)")
    for mode in a.modes.split(","):
        im, om, lc, sc, dst, *rest = maps[mode]
        src = rest[0] if rest else x
        om = om or im
        for c_log in [int(v) for v in a.clogs.split(",")]:
            for copy in ([False, True] if a.copy else [False]):
                ry = (0, 0, 0, 0) if copy else (-1, 0, 0, 10)
                if copy and lc != sc:
                    continue
                try:
                    ob.fft_launch_raw(src, dst, n, nb, im, om, bits=bits, c_log=c_log, load_cfast=lc, store_cfast=sc, ry=ry, repeat=2)
                    ms = ob.fft_launch_raw(src, dst, n, nb, im, om, bits=bits, c_log=c_log, load_cfast=lc, store_cfast=sc, ry=ry, repeat=rep)
                    print(f"  {mode:8s} {'copy' if copy else 'fft '} c_log={c_log:2d}: {ms:8.3f} ms  {gb / ms * 1e3:8.1f} GB/s", flush=True)
                except ob.OfftError as e:
                    print(f"  {mode:3s} c_log={c_log:2d}: skipped ({str(e)[-70:]})")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rep):
        y.copy_(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / rep
    print(f"  torch copy_: {ms:8.3f} ms  {gb / ms * 1e3:8.1f} GB/s")
    if a.plan and rows == n * n:
        for S in (1, 0):
            plan = ob.Plan(n, n, n, is_notest=1, custom={ob.P.P1: 1, ob.P.S: S})
            plan.set_stage_timing(True)
            for _ in range(3):
                plan.execute(x)
            print(f"  plan S={S}: {plan.last_ms:.3f} ms  stages {plan.stage_ms()}  GFLOP/s {5 * n ** 3 * 3 * (n.bit_length() - 1) / plan.last_ms / 1e6:.0f}")
            plan.fin()
    ob.world_fin()


if __name__ == "__main__":
    main()
