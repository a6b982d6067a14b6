"""Multi-GPU frame: one process per GPU, a BVH replica on each, image rows dealt out in stripes.

Replaces the reference's only parallel construct, ``tbb::parallel_for`` over image rows
(main.cpp:329-331): rows are independent units (own RNG streams, disjoint pixels, read-only
scene), so there is NO collective on the data path while rendering.  Stripe ``k`` (rows
``[k*stripe, (k+1)*stripe)``) belongs to rank ``k % world``; per-pixel RNG seeds depend only on
the pixel index, so the frame is identical for any world size.  The one exchange step is the
frame gather at the end: every rank's packed stripes -> rank 0 (``torch.distributed``, NCCL on
GPUs), where a CUDA kernel (tmpt_unpack_stripes) scatters them into the frame; plus an 8-byte
sum of the ray counters.

torch is plumbing here (device buffers, streams, process group); the pixels come from
libtmpt.so.
"""
from __future__ import annotations

import numpy as np

DEFAULT_STRIPE_ROWS = 4  # one tile row of the render kernel's 8x4 pixel tiles


def stripe_plan(height: int, stripe: int, world: int):
    """Rows owned by each rank and the padded per-rank row count of the gather buffer."""
    stripes = (height + stripe - 1) // stripe
    rows = [0] * world
    for k in range(stripes):
        rows[k % world] += min(stripe, height - k * stripe)
    return rows, max(rows)


def owned_rows(height: int, stripe: int, rank: int, world: int) -> np.ndarray:
    """Global row index of each packed row of `rank`, in packed order."""
    ys = [y for y in range(height) if (y // stripe) % world == rank]
    return np.asarray(ys, np.int64)


def gather_frame(packed, width: int, height: int, stripe: int, rank: int, world: int, group=None, unpack=None):
    """All ranks call this with their packed stripes (uint8 tensor [maxRows, width, 4], padded).

    Returns the full frame tensor [height, width, 4] on rank 0 and None elsewhere.  `unpack`
    is for the CPU (gloo) tests only; on CUDA tensors the product kernel does the scatter and
    on CPU tensors without `unpack` this raises -- there is no CPU path."""
    import torch
    import torch.distributed as dist

    rows, max_rows = stripe_plan(height, stripe, world)
    assert packed.shape == (max_rows, width, 4) and packed.dtype == torch.uint8
    if world == 1:
        gathered = packed.unsqueeze(0)
    else:
        gathered = torch.empty((world, max_rows, width, 4), dtype=torch.uint8, device=packed.device) if rank == 0 else None
        if packed.is_cuda:
            # NCCL gather = grouped send/recv to rank 0 over NVLink
            dist.gather(packed, list(gathered.unbind(0)) if rank == 0 else None, dst=0, group=group)
        else:
            dist.gather(packed, [gathered[r] for r in range(world)] if rank == 0 else None, dst=0, group=group)
    if rank != 0:
        return None
    if unpack is not None:
        return unpack(gathered, width, height, stripe, world)
    if not packed.is_cuda:
        raise RuntimeError("gather_frame: CPU tensors and no CUDA device -- this package has no CPU path")
    from . import _check, lib
    frame = torch.empty((height, width, 4), dtype=torch.uint8, device=packed.device)
    stream = torch.cuda.current_stream(packed.device).cuda_stream
    _check(lib().tmpt_unpack_stripes(gathered.data_ptr(), width, height, stripe, world, packed.device.index or 0, frame.data_ptr(), stream))
    return frame


def render_frame(scene, camera, width: int, height: int, spp: int, rank: int, world: int, stripe: int = DEFAULT_STRIPE_ROWS,
                 group=None, device=None):
    """One frame across `world` ranks.  Returns (frame uint8 CUDA tensor on rank 0 else None, local ray count tensor).

    Asynchronous on torch's current stream of `device` apart from the collective's own
    synchronisation; the caller brackets it with CUDA events."""
    import torch

    device = device or torch.device("cuda", scene.device)
    rows, max_rows = stripe_plan(height, stripe, world)
    packed = torch.empty((max_rows, width, 4), dtype=torch.uint8, device=device)
    rays = torch.zeros(1, dtype=torch.int64, device=device)
    stream = torch.cuda.current_stream(device).cuda_stream
    if stream == 0:
        # NULL means "the scene's own stream" to the C ABI, which would not be ordered with torch's work
        raise RuntimeError("render_frame: run inside `with torch.cuda.stream(torch.cuda.Stream())` (non-default stream)")
    scene.render_stripes(camera, width, height, spp, stripe, rank, world, packed.data_ptr(), rays.data_ptr(), stream=stream)
    frame = gather_frame(packed, width, height, stripe, rank, world, group=group)
    return frame, rays
