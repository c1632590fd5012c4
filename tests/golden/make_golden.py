"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/ref_dump,
built from /root/reference by `make -C oracle ref`) on forked host ranks.  Run in the build
container only; the fixtures are committed because /root/reference does not travel.

Each fixture holds the run's arguments, every rank's `struct _offt_comm` box, the final
tunables, and every rank's raw in-place output array (complex128).  Inputs are not stored:
they are `oracle.grid_values(seed, Nx, Ny, Nz)`.

    python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402

P1, T1, W1, T2, W2, V, S, RY = O.P1, O.T1, O.W1, O.T2, O.W2, O.V, O.S, O.RY

# name: (N, p, is_oned, is_equalxy, custom params)
CASES = {
    "pencil_2x2_16": ((16, 16, 16), 4, 0, 0, {P1: 2}),
    "pencil_2x2_16_stride": ((16, 16, 16), 4, 0, 0, {P1: 2, S: 1}),
    "pencil_2x2_16_equalxy": ((16, 16, 16), 4, 0, 1, {P1: 2}),
    "pencil_2x4_aniso": ((32, 64, 16), 8, 0, 0, {P1: 2, T1: 4, T2: 2}),
    "pencil_4x2_aniso_stride": ((32, 64, 16), 8, 0, 0, {P1: 4, S: 1, RY: 3}),
    "slab_px1_16_8_32": ((16, 8, 32), 4, 1, 0, {P1: 4}),
    "slab_px1_16_8_32_stride": ((16, 8, 32), 4, 1, 0, {P1: 4, S: 1}),
    "slab_1xp_16_8_32": ((16, 8, 32), 4, 1, 0, {P1: 1}),
    "slab_1xp_16_8_32_stride": ((16, 8, 32), 4, 1, 0, {P1: 1, S: 1}),
    "single_rank_16_8_32": ((16, 8, 32), 1, 0, 0, {P1: 1}),
    "single_rank_oned_stride": ((16, 8, 32), 1, 1, 0, {P1: 1, S: 1}),
    "ragged_tiles_2x2": ((32, 32, 32), 4, 0, 0, {P1: 2, T1: 3, T2: 5, W1: 1, W2: 3}),
    "uneven_12_10_9_a2av": ((12, 10, 9), 4, 0, 0, {P1: 2, V: 3}),
    "uneven_12_10_9_padded_stride": ((12, 10, 9), 4, 0, 0, {P1: 2, V: 0, S: 1}),
    "uneven_3ranks_slab": ((12, 10, 9), 3, 1, 0, {P1: 3, V: 3, T2: 2}),
    "uneven_6ranks": ((20, 12, 18), 6, 0, 0, {P1: 3, V: 3, T1: 3, T2: 4}),
}
SEED = 20161
# real-to-complex plans (is_r2c = 1): the real parts of the same seeded grid, fixtures in tests/golden/r2c/
R2C_CASES = {
    "r2c_single_12_10_9": ((12, 10, 9), 1, 0, 0, {P1: 1}),
    "r2c_single_stride_16_8_32": ((16, 8, 32), 1, 1, 0, {P1: 1, S: 1}),
    "r2c_slab_px1_16_8_32": ((16, 8, 32), 4, 1, 0, {P1: 4}),
    "r2c_slab_1xp_stride_16_8_32": ((16, 8, 32), 4, 1, 0, {P1: 1, S: 1}),
    "r2c_pencil_2x2_12_10_18_a2av": ((12, 10, 18), 4, 0, 0, {P1: 2, V: 3}),
    "r2c_pencil_2x4_32_64_16": ((32, 64, 16), 8, 0, 0, {P1: 2, T1: 4, T2: 2}),
}

if __name__ == "__main__":
    if not O.have_reference():
        sys.exit("oracle/_ref/ref_dump missing: run `make -C oracle ref` where /root/reference exists")
    out_dir = Path(__file__).resolve().parent
    for name, (N, p, oned, eq, params) in CASES.items():
        boxes, _ = O.run_reference(*N, p, seed=SEED, is_oned=oned, is_equalxy=eq, params=params)
        desc = np.array([[b.p1, b.p2, *b.istart, *b.isize, *b.istride, *b.ostart, *b.osize, *b.ostride, b.alloc]
                         for b in boxes], dtype=np.int64)
        np.savez_compressed(
            out_dir / f"{name}.npz",
            N=np.array(N), p=p, is_oned=oned, is_equalxy=eq, seed=SEED,
            custom=np.array(sorted(params.items()), dtype=np.int64).reshape(-1, 2),
            params=np.array(boxes[0].params, dtype=np.int32), desc=desc,
            data=np.concatenate([b.data for b in boxes]))
        err = O.rel_l2(O.gather_output(boxes), np.fft.fftn(O.grid_values(SEED, *N)))
        print(f"{name}: p={p} N={N} rel-L2 vs numpy {err:.2e}")
    (out_dir / "r2c").mkdir(exist_ok=True)
    for name, (N, p, oned, eq, params) in R2C_CASES.items():
        boxes, _ = O.run_reference(*N, p, seed=SEED, is_oned=oned, is_equalxy=eq, params=params, is_r2c=1)
        desc = np.array([[b.p1, b.p2, *b.istart, *b.isize, *b.istride, *b.ostart, *b.osize, *b.ostride, b.alloc]
                         for b in boxes], dtype=np.int64)
        np.savez_compressed(
            out_dir / "r2c" / f"{name}.npz",
            N=np.array(N), p=p, is_oned=oned, is_equalxy=eq, seed=SEED, is_r2c=1,
            custom=np.array(sorted(params.items()), dtype=np.int64).reshape(-1, 2),
            params=np.array(boxes[0].params, dtype=np.int32), desc=desc,
            data=np.concatenate([b.data for b in boxes]))
        err = O.rel_l2(O.gather_output_r2c(boxes), np.fft.rfftn(O.grid_values(SEED, *N).real))
        print(f"{name}: p={p} N={N} r2c rel-L2 vs numpy rfftn {err:.2e}")
