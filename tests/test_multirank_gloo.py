"""World-size-2 (and 3) CPU test of the multi-GPU host logic over gloo: stripe ownership,
packed layout and the frame gather (toymeshpathtracer_b200/multigpu.py).  The pixels are
synthetic here -- rendering needs a GPU -- and the scatter is checked with a numpy
restatement of tmpt_unpack_stripes (the CUDA kernel itself is covered by the gpu tests)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import toymeshpathtracer_b200 as tm
from toymeshpathtracer_b200 import multigpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _numpy_unpack(gathered, width, height, stripe, world):
    g = gathered.numpy()
    frame = np.zeros((height, width, 4), np.uint8)
    if stripe == 0:  # tile interleave: rank (tx + ty) % world, local tile tx // world (k_unpack_stripes)
        owner, xl = multigpu.tile_owner_maps(width, height, world)
        ys = np.arange(height)[:, None].repeat(width, 1)
        return torch.from_numpy(g[owner, ys, xl])
    for r in range(world):
        ys = multigpu.owned_rows(height, stripe, r, world)
        frame[ys] = g[r, : len(ys)]
    return torch.from_numpy(frame)


def _pixel(y, x):  # a value every pixel can be checked by
    return np.stack([y % 251, x % 241, (y * 7 + x * 3) % 239, np.full_like(y, 255)], -1).astype(np.uint8)


def _worker(rank, world, port, width, height, stripe, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rows, max_rows = multigpu.stripe_plan(height, stripe, world)
        lw = multigpu.local_width(width, stripe, world)
        assert rows[rank] == tm.stripe_rows(height, stripe, rank, world) and lw == tm.local_width(width, stripe, world)  # host plan == C ABI
        ys = multigpu.owned_rows(height, stripe, rank, world)
        assert len(ys) == rows[rank]
        packed = np.zeros((max_rows, lw, 4), np.uint8)
        if stripe == 0:  # what this rank's render kernel would write: its tiles, local tile ltx of tile row ty = frame tile ((rank - ty) % world) + ltx * world
            yy, xl = np.meshgrid(np.arange(height), np.arange(lw), indexing="ij")
            xx = (((rank - yy // 4) % world) + (xl // 8) * world) * 8 + xl % 8
            packed[:] = np.where((xx < width)[..., None], _pixel(yy, np.minimum(xx, width - 1)), 0)
        else:
            yy, xx = np.meshgrid(ys, np.arange(width), indexing="ij")
            packed[: len(ys)] = _pixel(yy, xx)
        frame = multigpu.gather_frame(torch.from_numpy(packed), width, height, stripe, rank, world, unpack=_numpy_unpack)
        rays = torch.tensor([1000 + rank], dtype=torch.int64)
        dist.all_reduce(rays)
        if rank == 0:
            yy, xx = np.meshgrid(np.arange(height), np.arange(width), indexing="ij")
            ok = bool((frame.numpy() == _pixel(yy, xx)).all()) and int(rays) == sum(1000 + r for r in range(world))
            out.put(ok)
        else:
            assert frame is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,width,height,stripe", [(2, 64, 36, 4), (2, 33, 13, 4), (3, 16, 50, 8), (2, 64, 36, 0), (3, 50, 23, 0), (2, 7, 3, 0)])
def test_stripe_gather_over_gloo(world, width, height, stripe):
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, width, height, stripe, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get() is True


def test_cpu_tensors_without_checker_fail_loudly():
    packed = torch.zeros((8, 4, 4), dtype=torch.uint8)
    with pytest.raises(RuntimeError, match="no CPU path"):
        multigpu.gather_frame(packed, 4, 8, 4, 0, 1)
