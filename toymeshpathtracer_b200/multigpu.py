"""Multi-GPU frame: one process per GPU, a BVH replica on each, the image dealt out in 8x4-pixel tiles.

Replaces the reference's only parallel construct, ``tbb::parallel_for`` over image rows
(main.cpp:329-331): pixels are independent units (own RNG streams, disjoint bytes, read-only
scene), so there is NO collective on the data path while rendering.  Default partition
(``stripe = 0``, "tile interleave"): tile ``(tx, ty)`` belongs to rank ``(tx + ty) % world`` --
every rank owns exactly 1/world of the tiles, spread evenly over the frame (balanced in count and
in cost).  ``stripe > 0`` selects row stripes instead: stripe ``k`` (rows ``[k*stripe,
(k+1)*stripe)``) belongs to rank ``k % world``.  Per-pixel RNG seeds depend only on the pixel
index, so the frame is identical for any world size and either partition.

By default there is no gather either: every rank's render kernel stores its pixels straight
into rank 0's frame over NVLink (``PeerFrame``).  The fallback exchange step is a gather of
every rank's packed share -> rank 0 (``torch.distributed``, NCCL on GPUs), where a CUDA kernel
(tmpt_unpack_stripes) scatters them into the frame; plus an 8-byte sum of the ray counters.

torch is plumbing here (device buffers, streams, process group); the pixels come from
libtmpt.so.
"""
from __future__ import annotations

import numpy as np

DEFAULT_STRIPE_ROWS = 0  # tile interleave (row stripes of 4 rows: two of eight ranks get 33 instead of 34 stripes of 1080 rows)


def local_width(width: int, stripe: int, world: int) -> int:
    """Pixels per row of a rank's packed share (== tmpt_local_width)."""
    return width if stripe > 0 else (((width + 7) // 8 + world - 1) // world) * 8


def tile_owner_maps(width: int, height: int, world: int):
    """Tile interleave, restated in numpy for the CPU tests: per frame pixel the owning rank and its local x."""
    ys, xs = np.meshgrid(np.arange(height), np.arange(width), indexing="ij")
    tx, ty = xs // 8, ys // 4
    return (tx + ty) % world, (tx // world) * 8 + (xs % 8)


def stripe_plan(height: int, stripe: int, world: int):
    """Rows of each rank's packed share and the padded per-rank row count of the gather buffer."""
    if stripe == 0:
        return [height] * world, height
    stripes = (height + stripe - 1) // stripe
    rows = [0] * world
    for k in range(stripes):
        rows[k % world] += min(stripe, height - k * stripe)
    return rows, max(rows)


def owned_rows(height: int, stripe: int, rank: int, world: int) -> np.ndarray:
    """Global row index of each packed row of `rank`, in packed order."""
    if stripe == 0:
        return np.arange(height, dtype=np.int64)
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
    lw = local_width(width, stripe, world)
    assert packed.shape == (max_rows, lw, 4) and packed.dtype == torch.uint8
    if world == 1:
        gathered = packed.unsqueeze(0)
    else:
        gathered = torch.empty((world, max_rows, lw, 4), dtype=torch.uint8, device=packed.device) if rank == 0 else None
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


class _DevArray:
    """Zero-copy view of library-owned device memory for torch (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "|u1", "data": (int(ptr), False), "version": 3, "strides": None}


class PeerFrame:
    """Rank 0's full-size frame, mapped into every rank through CUDA IPC (tmpt_frame_alloc / tmpt_frame_open).

    With it the gather is fused into the render kernel: each rank's pixels are plain stores to
    rank 0's HBM over NVLink / NVSwitch (`peerFrame` of tmpt_render_stripes); what remains of
    the exchange step is a one-element all-reduce enqueued behind every rank's kernel, which
    orders "all pixels of frame N are stored" before rank 0's stream continues.

    The frame is DOUBLE-BUFFERED: consecutive frames alternate between two buffers.  Without that a
    rank could start storing frame N+1 while rank 0 is still reading frame N (its all-reduce only
    says that every rank has finished RENDERING frame N).  With two buffers, buffer A is written
    again by frame N+2, which no rank starts before the all-reduce of frame N+1 has completed --
    and on rank 0 that all-reduce sits on the stream BEHIND whatever consumed frame N.  The
    requirement on the caller: consume a frame on the stream it was rendered on (or synchronise
    with it) before enqueueing the frame after the next."""

    def __init__(self, width: int, height: int, rank: int, world: int, device_index: int, group=None):
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _check, lib
        self.rank, self.world, self.dev = rank, world, device_index
        self.shape = (height, width, 4)
        self.frame_bytes = (width * height * 4 + 255) // 256 * 256
        self.index = 0  # buffer of the next frame
        self.ptr = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        err = None
        if rank == 0:
            try:
                _check(lib().tmpt_frame_alloc(device_index, 2 * self.frame_bytes, C.byref(self.ptr), handle))
            except Exception as e:  # noqa: BLE001 -- still reach the broadcast so the other ranks do not hang
                err = e
        box = [bytes(handle) if err is None else b""]
        if world > 1:
            dist.broadcast_object_list(box, src=0, group=group)
        if err is not None:
            raise err
        if not box[0]:
            raise RuntimeError("rank 0 could not allocate the shared frame")
        if rank != 0:
            h = (C.c_ubyte * 64).from_buffer_copy(box[0])
            _check(lib().tmpt_frame_open(device_index, h, C.byref(self.ptr)))
        self.tensors = [torch.as_tensor(_DevArray(self.ptr.value + k * self.frame_bytes, self.shape), device=torch.device("cuda", device_index))
                        for k in range(2)] if rank == 0 else None

    def next_buffer(self):
        """-> (device address every rank stores the coming frame to, rank 0's tensor view of it); then flips."""
        k = self.index
        self.index ^= 1
        return self.ptr.value + k * self.frame_bytes, (self.tensors[k] if self.tensors else None)

    def close(self):
        from . import lib
        if self.ptr:
            (lib().tmpt_frame_free if self.rank == 0 else lib().tmpt_frame_close)(self.dev, self.ptr)
            self.ptr = None


def render_frame(scene, camera, width: int, height: int, spp: int, rank: int, world: int, stripe: int = DEFAULT_STRIPE_ROWS,
                 group=None, device=None, peer: "PeerFrame | None" = None):
    """One frame across `world` ranks.  Returns (frame uint8 CUDA tensor on rank 0 else None, local ray count tensor).

    Asynchronous on torch's current stream of `device` apart from the collective's own
    synchronisation; the caller brackets it with CUDA events.  With `peer` the pixels go
    straight into rank 0's frame (no gather); a one-element all-reduce enqueued behind every
    rank's render kernel orders the frame on rank 0."""
    import torch

    device = device or torch.device("cuda", scene.device)
    if peer is not None:
        import torch.distributed as dist
        rays = torch.zeros(1, dtype=torch.int64, device=device)
        stream = torch.cuda.current_stream(device).cuda_stream
        if stream == 0:
            raise RuntimeError("render_frame: run inside `with torch.cuda.stream(torch.cuda.Stream())` (non-default stream)")
        frame_ptr, frame = peer.next_buffer()  # (every rank flips in step: one call per frame on each)
        scene.render_stripes(camera, width, height, spp, stripe, rank, world, 0, rays.data_ptr(), peer_frame_ptr=frame_ptr, stream=stream)
        if world > 1:
            token = torch.ones(1, dtype=torch.int32, device=device)
            dist.all_reduce(token, group=group)  # enqueued behind the render kernel on every rank
        return frame, rays
    rows, max_rows = stripe_plan(height, stripe, world)
    packed = torch.empty((max_rows, local_width(width, stripe, world), 4), dtype=torch.uint8, device=device)
    rays = torch.zeros(1, dtype=torch.int64, device=device)
    stream = torch.cuda.current_stream(device).cuda_stream
    if stream == 0:
        # NULL means "the scene's own stream" to the C ABI, which would not be ordered with torch's work
        raise RuntimeError("render_frame: run inside `with torch.cuda.stream(torch.cuda.Stream())` (non-default stream)")
    scene.render_stripes(camera, width, height, spp, stripe, rank, world, packed.data_ptr(), rays.data_ptr(), stream=stream)
    frame = gather_frame(packed, width, height, stripe, rank, world, group=group)
    return frame, rays
