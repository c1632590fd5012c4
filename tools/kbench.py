"""Per-pass timing of the three local passes of an N^3 complex FFT on one GPU (device-resident,
CUDA events around `repeat` back-to-back launches).  Prints GB/s = 2*B*N^3 / t per pass."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import offt_b200 as ob  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    bits = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    rep = 10
    ob.world_fin(); ob.world_init_local(1, 0)
    dt = torch.complex128 if bits == 64 else torch.complex64
    a = torch.view_as_complex(torch.rand((n, n, n, 2), device="cuda", dtype=torch.float64 if bits == 64 else torch.float32)).contiguous()
    b = torch.empty_like(a)
    esz = 16 if bits == 64 else 8
    gb = 2 * esz * n ** 3 / 1e9
    N2, N3 = n * n, n * n * n
    z_map = [0, 0, 0, 1, n, n, n, N2, 0]
    y_map = [0, 0, 0, n, n, 1, n, N2, 0]
    x_map = [0, 0, 0, N2, n, 1, n, n, 0]
    xt_map = [0, 0, 0, 1, n, N2, n, n, 0]       # z-y-x output
    print(f"N={n}^3 bits={bits}: {gb:.2f} GB per pass (algorithmic)")
    for name, im, om, lc, sc, dst in (("z contiguous", z_map, z_map, 0, 0, a), ("y strided", y_map, y_map, 1, 1, a),
                                      ("x strided", x_map, x_map, 1, 1, a), ("x transposing", x_map, xt_map, 1, 0, b)):
        for c_log in (-1, 0, 1, 2, 3, 4):
            try:
                ob.fft_launch_raw(a, dst, n, N2, im, om, bits=bits, c_log=c_log, load_cfast=lc, store_cfast=sc, repeat=2)
                ms = ob.fft_launch_raw(a, dst, n, N2, im, om, bits=bits, c_log=c_log, load_cfast=lc, store_cfast=sc, repeat=rep)
                print(f"  {name:14s} c_log={c_log:2d}: {ms:8.3f} ms  {gb / ms * 1e3:8.1f} GB/s")
            except ob.OfftError as e:
                print(f"  {name:14s} c_log={c_log:2d}: skipped ({str(e)[-60:]})")
    # plain copy for reference
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rep):
        b.copy_(a)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / rep
    print(f"  torch copy_: {ms:8.3f} ms  {gb / ms * 1e3:8.1f} GB/s")
    # whole plan
    for S in (1, 0):
        plan = ob.Plan(n, n, n, is_notest=1, custom={ob.P.P1: 1, ob.P.S: S})
        plan.set_stage_timing(True)
        for _ in range(3):
            plan.execute(a)
        print(f"  plan S={S}: {plan.last_ms:.3f} ms  stages {plan.stage_ms()}  GFLOP/s {5 * N3 * 3 * (n.bit_length() - 1) / plan.last_ms / 1e6:.0f}")
        plan.fin()
    ob.world_fin()


if __name__ == "__main__":
    main()
