"""Host-side views of the rank-local in-place arrays through the layout descriptor.

`struct _offt_comm` (include/offt.h; offt.h:102-142 of the reference) tells the caller where its data lives:
the input box `istart/isize/istride` before a forward transform (run-fft.c:49-57 fills it that way) and the
output box `ostart/osize/ostride` after it (run-fft.c:477-478 reads it that way).  These helpers turn a box
(the dict of `offt_b200.comm_box` / `Plan.box()`) and a flat complex array (numpy or torch, host or device)
into a strided 3-D view, so that callers - bench.py's parity gate, tools, user code - can fill and read
distributed grids without re-deriving the index expressions.  No transform arithmetic lives here.
"""
from __future__ import annotations

import numpy as np


def _view(flat, size, stride):
    if hasattr(flat, "as_strided"):            # torch tensor (host or device)
        return flat.as_strided(tuple(int(s) for s in size), tuple(int(s) for s in stride))
    it = flat.itemsize
    return np.lib.stride_tricks.as_strided(flat, shape=tuple(int(s) for s in size), strides=tuple(int(s) * it for s in stride))


def input_view(box: dict, flat):
    """[isize] view of this rank's part of the global input grid inside its flat in-place array"""
    return _view(flat, box["isize"], box["istride"])


def output_view(box: dict, flat):
    """[osize] view of this rank's part of the global spectrum inside its flat in-place array"""
    return _view(flat, box["osize"], box["ostride"])


def _slices(start, size):
    return tuple(slice(int(a), int(a) + int(n)) for a, n in zip(start, size))


def scatter_input(box: dict, grid: np.ndarray, alloc: int, dtype=None) -> np.ndarray:
    """flat array of `alloc` elements holding `grid`'s input box of this rank (zeros elsewhere)"""
    flat = np.zeros(int(alloc), dtype=dtype or grid.dtype)
    input_view(box, flat)[...] = grid[_slices(box["istart"], box["isize"])]
    return flat


def gather_output(boxes: list, arrays: list, N, dtype=np.complex128) -> np.ndarray:
    """global [Nx, Ny, Nz] spectrum assembled from every rank's array through ostart/osize/ostride"""
    out = np.full(tuple(N), np.nan + 0j, dtype=dtype)
    for box, a in zip(boxes, arrays):
        if min(box["osize"]) > 0:
            out[_slices(box["ostart"], box["osize"])] = output_view(box, np.asarray(a))
    return out


def gather_input(boxes: list, arrays: list, N, dtype=np.complex128) -> np.ndarray:
    """global grid re-assembled through istart/isize/istride (what a backward transform returns)"""
    out = np.full(tuple(N), np.nan + 0j, dtype=dtype)
    for box, a in zip(boxes, arrays):
        if min(box["isize"]) > 0:
            out[_slices(box["istart"], box["isize"])] = input_view(box, np.asarray(a))
    return out
