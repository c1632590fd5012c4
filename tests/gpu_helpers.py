"""helpers for the -m gpu parity tests: drive the C-ABI library through its ctypes binding"""
from contextlib import contextmanager

import numpy as np

from oracle import oracle as O


@contextmanager
def local_world(size):
    import offt_b200 as ob
    ob.world_fin()
    ob.world_init_local(size, 0)
    try:
        yield ob
    finally:
        ob.world_fin()


def box_of(plan, N, p):
    c = plan.comm
    return O.RankBox(p=p, rank=plan.rank, N=tuple(N), p1=c.p1, p2=c.p2, istart=tuple(c.istart), isize=tuple(c.isize),
                     istride=tuple(c.istride), ostart=tuple(c.ostart), osize=tuple(c.osize), ostride=tuple(c.ostride),
                     alloc=plan.alloc_elems, params=plan.params)


def gpu_forward(grid, p, custom, is_oned=0, is_equalxy=0, bits=64, host_arrays=False, inverse_too=False, is_r2c=0):
    """forward 3-D FFT of the global `grid` on p emulated ranks of one GPU through the C API.
    Returns (boxes with .data = each rank's in-place array after the transform, plans' launch count,
    optionally the arrays after a following backward transform)."""
    import torch
    N = grid.shape
    cdt = np.complex128 if bits == 64 else np.complex64
    tdt = torch.complex128 if bits == 64 else torch.complex64
    with local_world(p) as ob:
        ob.set_default_precision(bits)
        try:
            plans = [ob.Plan(*N, is_oned=is_oned, is_equalxy=is_equalxy, is_notest=1, custom=custom, rank=r, is_r2c=is_r2c) for r in range(p)]
            boxes = [box_of(pl, N, p) for pl in plans]
            if is_r2c:   # real input in the in-place r2c layout (run-fft.c:53-55), handed over as the complex view of the doubles
                host = [np.ascontiguousarray(O.scatter_input_r2c(b, np.asarray(grid.real, dtype=np.float64))) for b in boxes]
                if bits == 32:
                    host = [h.view(np.float64).astype(np.float32).view(np.complex64) for h in host]
            else:
                host = [np.ascontiguousarray(O.scatter_input(b, grid.astype(cdt))) for b in boxes]
            if host_arrays:
                arrays = host
            else:
                arrays = [torch.from_numpy(h).to("cuda") for h in host]
            if p == 1:
                plans[0].execute(arrays[0])
            else:
                ob.execute_group(plans, arrays)
            launches = sum(pl.last_launches for pl in plans)
            for b, a in zip(boxes, arrays):
                b.data = a.copy() if host_arrays else a.cpu().numpy()
            back = None
            if inverse_too:
                if p == 1:
                    plans[0].execute_inverse(arrays[0])
                else:
                    ob.execute_group(plans, arrays, inverse=True)
                back = [a.copy() if host_arrays else a.cpu().numpy() for a in arrays]
            for pl in plans:
                pl.fin()
        finally:
            ob.set_default_precision(64)
    return boxes, launches, back


def gather_input(boxes, arrays, dtype=np.complex128):
    """global grid re-assembled through istart/isize/istride"""
    Nx, Ny, Nz = boxes[0].N
    out = np.full((Nx, Ny, Nz), np.nan + 0j, dtype=dtype)
    for b, a in zip(boxes, arrays):
        sx, sy, sz = b.isize
        ix = np.arange(sx)[:, None, None] * b.istride[0]
        iy = np.arange(sy)[None, :, None] * b.istride[1]
        iz = np.arange(sz)[None, None, :] * b.istride[2]
        out[b.istart[0]:b.istart[0] + sx, b.istart[1]:b.istart[1] + sy, b.istart[2]:b.istart[2] + sz] = \
            a[(ix + iy + iz).ravel()].reshape(sx, sy, sz)
    return out
