"""CPU, two processes over gloo: the host-side bookkeeping of the multi-rank path through the C-ABI library -
rank boxes, exchange block sizes, default tunables - without a GPU.  Each rank lays its input box out exactly as
the GPU path expects, the phase-2 exchange is carried out with torch.distributed.all_to_all on blocks of
offtb_exchange_block_elems, and the unpacked result must equal the oracle's layout for that rank.
(The device kernels themselves are covered by the -m gpu tests; nothing here computes an FFT in the product.)"""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, N, ret):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    import offt_b200 as ob
    from oracle import oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        Nx, Ny, Nz = N
        p1 = world                                     # slab p x 1 (x slabs -> y slabs), phase 2 only
        box = ob.comm_box(Nx, Ny, Nz, world, p1, rank, S=1)
        orc = O.Oracle()
        want_box = orc.box(Nx, Ny, Nz, world, p1, rank, S=1)
        for k in "istart isize istride ostart osize ostride".split():
            assert tuple(box[k]) == tuple(getattr(want_box, k)), k
        assert ob.alloc_elems(Nx, Ny, Nz, world, p1) == want_box.alloc
        v = ob.params_default(Nx, Ny, Nz, world, 0, 1)
        assert v == orc.params_default(Nx, Ny, Nz, world, 0, 1)
        # the exchange of one whole "tile" (all z planes): block a of my send buffer goes to rank a.
        # Data = the global index of each point, so that the landing place of every point can be checked.
        # Blocks are padded to the ceilings M1 x M4 (offt-compute.c:3704); with an uneven division every rank fills only
        # its own m1 x m4(a) corner - the last N % p owners hold one item more (offt-compute.c:141-144).
        boxes = [ob.comm_box(Nx, Ny, Nz, world, p1, r, S=1) for r in range(world)]
        M1, M4, m1, m4 = box["M1"], box["M4"], box["m1"], box["m4"]
        myT = Nz
        blk = M1 * M4 * myT                           # offtb_exchange_block_elems(po, 2, myT)
        x0 = box["istart"][0]
        gx, gy, gz = np.meshgrid(np.arange(x0, x0 + m1), np.arange(Ny), np.arange(Nz), indexing="ij")
        gid = ((gx * Ny + gy) * Nz + gz).astype(np.float64)
        send = np.full((world, M1, M4, myT), -1.0)
        for a in range(world):                         # pack2, S=1 layout [x][y_local][z] (offt-compute.c:1758-1776)
            ya, na = boxes[a]["ostart"][1], boxes[a]["m4"]
            send[a, :m1, :na, :] = gid[:, ya:ya + na, :]
        recv = torch.empty(world * blk, dtype=torch.float64)
        dist.all_to_all_single(recv, torch.from_numpy(send.reshape(-1)))
        recv = recv.numpy().reshape(world, M1, M4, myT)  # [source][x_local][y_local][z]
        # unpack2: out[z + M3*y + M3*M4*x] (offt-compute.c:2447-2450) -> my output box (all x, my y block, all z)
        out = np.concatenate([recv[a, :boxes[a]["m1"], :m4, :] for a in range(world)], axis=0)
        assert sum(b["m1"] for b in boxes) == Nx and [b["istart"][0] for b in boxes] == list(np.cumsum([0] + [b["m1"] for b in boxes[:-1]]))
        y0 = box["ostart"][1]
        ex, ey, ez = np.meshgrid(np.arange(Nx), np.arange(y0, y0 + m4), np.arange(Nz), indexing="ij")
        assert np.array_equal(out, ((ex * Ny + ey) * Nz + ez).astype(np.float64))
        ret.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        ret.put((rank, f"{type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N", [(16, 8, 4), (32, 32, 8), (15, 9, 4), (7, 11, 3)])   # the last two divide unevenly over 2 ranks
def test_two_host_ranks_exchange_bookkeeping(N):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, N, ret)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(ret.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert got == [(0, "ok"), (1, "ok")], got
