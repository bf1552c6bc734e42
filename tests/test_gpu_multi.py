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
