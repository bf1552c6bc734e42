"""Host-side logic of the N>1 path, exercised with world_size 2 over gloo on CPU: band/frame
partitioning, the band gather that assembles one framebuffer, and the Stats reduction. The band
renders come from the CPU oracle (the same role the per-rank GPU renders play on a B200 box)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from retrofire_b200 import scenes, shard


def test_row_bands_partition_and_alignment():
    for h in (1, 31, 32, 33, 1080, 2160, 4320):
        for world in (1, 2, 3, 4, 8):
            bands = shard.row_bands(h, world)
            assert len(bands) == world and bands[0][0] == 0 and bands[-1][1] == h
            for (a, b), (c, d) in zip(bands, bands[1:]):
                assert b == c and a <= b
            sizes = [b - a for a, b in bands]
            if h % world == 0:
                assert len(set(sizes)) == 1
            else:
                assert all(a % 32 == 0 for a, _ in bands if a < h)
                assert max(sizes) - min(sizes) <= 32 + (h % 32 != 0) * 32


def test_frame_slices_cover_every_frame_once():
    for n in (1, 7, 32, 1024):
        for world in (1, 2, 8):
            seen = sorted(f for r in range(world) for f in shard.frame_slice(n, world, r))
            assert seen == list(range(n))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import rfo
        import retrofire_b200 as rf
        sc = scenes.random_soup(600, 256, 200, seed=9, lanes_kind="color3", big=True)
        bands = shard.row_bands(sc.h, world)
        tgt = rfo.HostTarget(sc.w, sc.h, sc.fmt, True)
        tgt.band = bands[rank]
        tgt.clear(sc.ctx.color_clear, sc.ctx.depth_clear)
        st = rf.Stats()
        for d in sc.draws:
            st += rfo.render(d, tgt)
        color = torch.from_numpy(tgt.color.view(np.int32))
        depth = torch.from_numpy(tgt.depth)
        shard.gather_bands(color, bands, rank)
        shard.gather_bands(depth, bands, rank)
        total = shard.reduce_stats([st.frags.i, st.frags.o])
        # every rank must now hold the full frame == the unsharded render
        full = rfo.HostTarget(sc.w, sc.h, sc.fmt, True)
        full.clear(sc.ctx.color_clear, sc.ctx.depth_clear)
        fs = rf.Stats()
        for d in sc.draws:
            fs += rfo.render(d, full)
        ok = (np.array_equal(color.numpy().view(np.uint32), full.color) and np.array_equal(depth.numpy().view(np.uint32), full.depth.view(np.uint32))
              and total == [fs.frags.i, fs.frags.o] and st.prims.o == fs.prims.o)
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_band_gather_assembles_the_unsharded_frame_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


class _FakeDev:
    """Stands in for rf.Device in the peer-exchange helpers: hands out fake handles, records what it was attached to, and
    reports a replay on the attempts listed in `replay_on` (a pass re-run after arena growth)."""

    def __init__(self, rank, replay_on=()):
        self.rank, self.replay_on, self._replays, self.attempt, self.attached = rank, set(replay_on), 0, 0, None

    def peer_export(self):
        return bytes([self.rank + 1]) * 64

    def peer_attach(self, world, rank, table):
        self.attached = (world, rank, list(table))

    def replays(self):
        return self._replays

    def sync(self):
        self.attempt += 1
        if self.attempt in self.replay_on:
            self._replays += 1


class _FakeFb:
    def __init__(self, rank, k):
        self.tag, self.attached = bytes([16 * (k + 1) + rank]) * 64, None

    def peer_export(self):
        return self.tag

    def peer_attach(self, world, rank, table):
        self.attached = list(table)


def _peer_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # handle exchange: every rank sees every rank's handles in rank order; root=0 blanks the entries of the others
        dev, fbs = _FakeDev(rank), [_FakeFb(rank, 0), _FakeFb(rank, 1)]
        shard.attach_peers(dev, fbs, rank, world)
        ok = dev.attached == (world, rank, [bytes([r + 1]) * 64 for r in range(world)])
        ok = ok and all(fb.attached == [bytes([16 * (k + 1) + r]) * 64 for r in range(world)] for k, fb in enumerate(fbs))
        dev2, fb2 = _FakeDev(rank), _FakeFb(rank, 0)
        shard.attach_peers(dev2, [fb2], rank, world, root=0)
        ok = ok and fb2.attached == [bytes([16 + r]) * 64 if r == 0 else bytes(64) for r in range(world)]
        # collective retry: only rank 1 replays, on its first attempt -> BOTH ranks render the frame twice
        dev3 = _FakeDev(rank, replay_on={1} if rank == 1 else ())
        frames = []
        tries = shard.render_frame_with_peers(dev3, lambda: frames.append(1))
        ok = ok and tries == 2 and len(frames) == 2
        dev4 = _FakeDev(rank)
        ok = ok and shard.render_frame_with_peers(dev4, lambda: None) == 1
        out[rank] = ok
    finally:
        dist.destroy_process_group()


def test_peer_handle_exchange_and_collective_retry_world2():
    """shard.attach_peers / render_frame_with_peers (the host side of csrc/rf_peer.cuh) over gloo with stand-in devices."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_peer_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}
