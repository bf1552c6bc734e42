"""Sort-first sharding across 2 GPUs with an NCCL gather of the finished row bands.
Needs >= 2 GPUs (skipped otherwise): run with `gpurun --gpus 2 -- python -m pytest tests -m gpu`."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import retrofire_b200 as rf
        from retrofire_b200 import scenes, shard
        from oracle import rfo
        sc = scenes.random_soup(3000, 1024, 768, seed=17, lanes_kind="lit", big=True)
        bands = shard.row_bands(sc.h, world)
        stream = torch.cuda.current_stream()
        dev = rf.Device(rank, stream=stream.cuda_stream or None)
        dev.set_row_band(*bands[rank])
        fb = dev.framebuf(sc.w, sc.h, sc.fmt, True)
        fb.clear(sc.ctx)
        for d in sc.draws:
            dev.render(d, fb)
        st = dev.stats(reset=True)
        color, depth = shard.target_tensor(fb), shard.target_tensor(fb, depth=True)
        shard.gather_bands(color, bands, rank)
        shard.gather_bands(depth, bands, rank)
        torch.cuda.synchronize()
        total = shard.reduce_stats([st.frags.i, st.frags.o], device="cuda")
        got_c, got_d = fb.download_color(), fb.download_depth()
        ref = rfo.HostTarget(sc.w, sc.h, sc.fmt, True)
        ref.clear(sc.ctx.color_clear, sc.ctx.depth_clear)
        rs = rf.Stats()
        for d in sc.draws:
            rs += rfo.render(d, ref)
        out[rank] = bool(np.array_equal(got_c, ref.host_color()) and np.array_equal(got_d.view(np.uint32), ref.depth.view(np.uint32))
                         and total == [rs.frags.i, rs.frags.o])
        dev.close()
    finally:
        dist.destroy_process_group()


def test_two_gpu_sort_first_with_nccl_gather():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}


def test_all_gpus_sort_first_with_nccl_gather():
    """The same frame cut into as many row bands as the box has GPUs (4 or 8): bands of 192 / 96 rows, gathered frame and summed
    Stats equal to the unsharded oracle frame."""
    import torch
    world = min(torch.cuda.device_count(), 8)
    if world < 4:
        pytest.skip("needs >= 4 GPUs (the 2-GPU case is the test above)")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {r: True for r in range(world)}


def _peer_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import retrofire_b200 as rf
        from retrofire_b200 import scenes, shard
        from oracle import rfo
        ok = True
        dev = rf.Device(rank)
        for sc in (scenes.random_soup(3000, 1024, 768, seed=17, lanes_kind="lit", big=True), scenes.bunny(subdiv=1, w=1024, h=768)):
            bands = shard.row_bands(sc.h, world)
            dev.set_row_band(*bands[rank])
            fb = dev.framebuf(sc.w, sc.h, sc.fmt, True)
            if sc.name.startswith("soup"):
                shard.attach_peers(dev, [fb], rank, world)
            else:   # a second target of the same ctx (barrier slots are attached once), gathered to rank 0 only
                table = [None] * world
                dist.all_gather_object(table, fb.peer_export())
                fb.peer_attach(world, rank, [t if r == 0 else bytes(64) for r, t in enumerate(table)])

            def frame():
                fb.clear(sc.ctx)
                for d in sc.draws:
                    dev.render(d, fb)

            tries = shard.render_frame_with_peers(dev, frame)     # first frame grows the arenas -> collective retry
            again = shard.render_frame_with_peers(dev, frame)
            got_c = fb.download_color()
            ref = rfo.HostTarget(sc.w, sc.h, sc.fmt, True)
            ref.clear(sc.ctx.color_clear, sc.ctx.depth_clear)
            for d in sc.draws:
                rfo.render(d, ref)
            want = ref.host_color()
            if sc.name.startswith("soup") or rank == 0:
                same = bool(np.array_equal(got_c, want))                  # the WHOLE frame (on every rank / on the root)
            else:
                y0, y1 = bands[rank]
                same = bool(np.array_equal(got_c[y0:y1], want[y0:y1]))    # a non-root rank holds its own band
            ok = ok and again == 1 and tries <= 3 and same
            dist.barrier()
        out[rank] = ok
        dev.close()
    finally:
        dist.destroy_process_group()


def test_two_gpu_sort_first_fused_peer_stores():
    """Sort-first with the exchange fused into the rasteriser: each rank's colour stores also go to the peer's framebuffer
    over NVLink (CUDA IPC), two cross-GPU barriers per frame, no gather. Every rank ends with the whole frame."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_peer_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}


def test_peer_stores_between_two_contexts_on_one_gpu(oracle):
    """The same mechanism inside ONE process (a host driving several contexts): peers named by raw device pointers. Two
    contexts on cuda:0 own the upper and the lower half of the frame; both end up with the complete frame."""
    import retrofire_b200 as rf
    from retrofire_b200 import scenes, shard
    world = 2
    for sc in (scenes.random_soup(2500, 800, 608, seed=23, lanes_kind="color3", big=True), scenes.sprites(2000, w=640, h=480)):
        devs = [rf.Device(0) for _ in range(world)]
        try:
            bands = shard.row_bands(sc.h, world)
            fbs = []
            for r, dev in enumerate(devs):
                dev.set_row_band(*bands[r])
                fbs.append(dev.framebuf(sc.w, sc.h, sc.fmt, True))
            flags = [dev.peer_export(ipc=False) for dev in devs]
            colors = [fb.peer_export(ipc=False) for fb in fbs]
            for r, dev in enumerate(devs):
                dev.peer_attach(world, r, flags)
                fbs[r].peer_attach(world, r, colors)
            for attempt in range(3):                      # the first frame may replay after arena growth
                before = [dev.replays() for dev in devs]
                for r, dev in enumerate(devs):            # queue and launch on BOTH contexts before waiting on either
                    fbs[r].clear(sc.ctx)
                    for d in sc.draws:
                        dev.render(d, fbs[r])
                    dev.flush()
                for dev in devs:
                    dev.sync()
                if all(dev.replays() == b for dev, b in zip(devs, before)):
                    break
            ref = oracle.HostTarget(sc.w, sc.h, sc.fmt, True)
            ref.clear(sc.ctx.color_clear, sc.ctx.depth_clear)
            want_stats = rf.Stats()
            for d in sc.draws:
                want_stats += oracle.render(d, ref)
            frags = [0, 0]
            for r, dev in enumerate(devs):
                assert np.array_equal(fbs[r].download_color(), ref.host_color()), f"{sc.name}: ctx {r} does not hold the whole frame"
                y0, y1 = bands[r]                          # depth stays sharded: each context holds its band
                assert np.array_equal(fbs[r].download_depth()[y0:y1].view(np.uint32), ref.depth[y0:y1].view(np.uint32))
        finally:
            for dev in devs:
                dev.close()


def test_all_gpus_sort_first_fused_peer_stores():
    """The fused exchange with every GPU of the box (4 or 8 ranks, each pushing its band into all the others)."""
    import torch
    world = min(torch.cuda.device_count(), 8)
    if world < 4:
        pytest.skip("needs >= 4 GPUs (the 2-GPU case is the test above)")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_peer_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {r: True for r in range(world)}
