"""Multi-GPU partitioning of the render() path on one box.

One process per GPU, torch.distributed for the plumbing. The path shards two ways, neither of
which needs a collective on the data path:

* frame/view sharding — frame f of a batch goes to rank f % world (or a contiguous block);
* sort-first sharding — geometry is replicated, rank r rasterises only the framebuffer rows of
  its band (`Device.set_row_band`), aligned to the 32-row tiles of the rasteriser.

The ONLY exchange is the final gather of finished bands into one framebuffer. Two forms:

* `attach_peers` (GPUs): the exchange is fused into the rasteriser — every colour store of a rank's band is also
  stored into the peers' framebuffers over NVLink (CUDA IPC peer memory), so after `Device.sync()` every rank holds
  the whole frame and nothing else is moved (`render_frame_with_peers` adds the collective retry an arena-growth
  replay needs);
* `gather_bands`: an all_gather of equal-sized (padded) row bands over NCCL/NVLink on GPUs, or gloo in the CPU tests.

Stats are summed with an all_reduce (`reduce_stats`).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

TILE = 32


def row_bands(h: int, world: int, tile: int = TILE) -> List[Tuple[int, int]]:
    """Contiguous row bands covering [0, h), one per rank: equal bands when h divides evenly, else whole tile rows."""
    if h % world == 0:
        # equal bands gather in place with one all_gather. They need not be tile aligned: a tile that straddles two bands is
        # rasterised by both owners, each producing only its own rows (the others are replaced by the gather).
        step = h // world
        return [(r * step, (r + 1) * step) for r in range(world)]
    tiles = (h + tile - 1) // tile
    per, extra = divmod(tiles, world)
    out, t0 = [], 0
    for r in range(world):
        t1 = t0 + per + (1 if r < extra else 0)
        out.append((min(t0 * tile, h), min(t1 * tile, h)))
        t0 = t1
    return out


def frame_slice(n_frames: int, world: int, rank: int) -> range:
    """Frames owned by `rank` under round-robin frame sharding."""
    return range(rank, n_frames, world)


class DeviceArray:
    """Zero-copy torch view of device memory owned by librf_b200 (rf_target_*_devptr)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(shape), "typestr": typestr, "version": 3, "strides": None}


def target_tensor(fb, depth: bool = False):
    """torch tensor (h, w) int32 / float32 aliasing a Framebuf's colour containers or depth buffer."""
    import torch
    ptr = fb.depth_devptr() if depth else fb.color_devptr()
    return torch.as_tensor(DeviceArray(ptr, (fb.h_px, fb.w), "<f4" if depth else "<i4"), device=f"cuda:{torch.cuda.current_device()}")


def gather_bands(buf, bands: Sequence[Tuple[int, int]], rank: int, group=None):
    """All-gather row bands of a (h, w) tensor: after the call every rank's `buf` holds every band.

    `buf` is this rank's full-size buffer in which only rows bands[rank] are valid. Bands are
    padded to the tallest band so that one all_gather_into_tensor moves everything.
    """
    import torch
    import torch.distributed as dist
    world = len(bands)
    h, w = buf.shape
    sizes = {b[1] - b[0] for b in bands}
    if len(sizes) == 1 and buf.is_cuda and buf.is_contiguous() and hasattr(dist, "all_gather_into_tensor"):
        # equal bands: in-place all-gather straight into the framebuffer (input is the rank's own slice of the output)
        y0, y1 = bands[rank]
        dist.all_gather_into_tensor(buf.view(-1), buf[y0:y1].reshape(-1), group=group)
        return buf
    rows = max(b[1] - b[0] for b in bands)
    send = torch.zeros((rows, w), dtype=buf.dtype, device=buf.device)
    y0, y1 = bands[rank]
    send[: y1 - y0] = buf[y0:y1]
    recv = torch.empty((world, rows, w), dtype=buf.dtype, device=buf.device)
    if hasattr(dist, "all_gather_into_tensor") and buf.is_cuda:
        dist.all_gather_into_tensor(recv.view(world * rows, w), send, group=group)
    else:
        parts = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(parts, send, group=group)
        recv = torch.stack(parts)
    for r, (a, b) in enumerate(bands):
        if r != rank and b > a:
            buf[a:b] = recv[r, : b - a]
    return buf


def reduce_stats(counters: Sequence[int], device=None, group=None) -> List[int]:
    """Sum Stats counters over ranks (frags.* are per-band under sort-first sharding)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(counters), dtype=torch.int64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return [int(x) for x in t]


# ---- sort-first with the exchange fused into the rasteriser (csrc/rf_peer.cuh) ------------------------------------
def attach_peers(dev, framebufs, rank: int, world: int, group=None, root: int = None) -> None:
    """One process per GPU: exchange CUDA IPC handles of the barrier slots and of every framebuffer's colour buffer with
    torch.distributed and attach them. Every rank must call this with the same number of framebufs, in the same order.
    root=None: every rank ends with the whole frame (all-gather semantics); root=r: only rank r collects it (gather)."""
    import torch.distributed as dist
    mine = [dev.peer_export()] + [fb.peer_export() for fb in framebufs]
    table = [None] * world
    dist.all_gather_object(table, mine, group=group)
    dev.peer_attach(world, rank, [t[0] for t in table])
    for i, fb in enumerate(framebufs):
        fb.peer_attach(world, rank, [t[1 + i] if root is None or r == root else bytes(64) for r, t in enumerate(table)])


def render_frame_with_peers(dev, draw_frame, group=None, max_tries: int = 4) -> int:
    """Run `draw_frame()` (clear + draws of ONE frame into peer-attached targets) and wait for it on every rank. If any
    rank had to re-run a pass after growing its arenas, peers may have seen the frame before those stores: the frame is
    rendered again (arenas only grow, so this settles). Returns the number of attempts."""
    import torch
    import torch.distributed as dist
    for attempt in range(1, max_tries + 1):
        before = dev.replays()
        draw_frame()
        dev.sync()
        grew = torch.tensor([dev.replays() - before], dtype=torch.int64, device="cuda" if dist.get_backend(group) == "nccl" else "cpu")
        dist.all_reduce(grew, op=dist.ReduceOp.MAX, group=group)
        if int(grew) == 0:
            return attempt
    raise RuntimeError("arenas kept growing")
