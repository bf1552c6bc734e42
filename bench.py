#!/usr/bin/env python
"""bench.py — headline benchmark of the render() hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload bunny|crates|sprites|small_tris]
                    [--frames F] [--sharding frames|tiles] [--exchange peer|peer-root|nccl]

Workload (BASELINE.json configs[1]): the Stanford bunny, 2x midpoint-subdivided to 79,488
triangles, Gouraud-shaded (VS_SOLIDS/FS_COLOR3F) with depth test at 1920x1080, Xrgb8888 + f32
depth. One "step" is one frame batch: F = 128 frames (theta = 2*pi*f/F; BASELINE config 5-ii is 1,024 frames
over 8 GPUs), each cleared and drawn into its own device-resident target — the frame-sharded batch of
SURVEY §8e. With N GPUs every rank renders its own F frames (weak scaling, no data-path collective).
`--sharding tiles` instead renders ONE large frame sort-first (row bands per rank; the band exchange is
fused into the rasteriser over NVLink peer memory, or `--exchange nccl` gathers the bands afterwards).

`value`   : Mfragments/s (Stats.frags.i per second) with geometry resident in HBM (rf_mesh).
`e2e`     : same metric through the reference-facing calls with HOST (page-locked) vertex/index buffers
            every frame (H2D inside the timed region) and the colour buffer of every frame downloaded into
            page-locked Buf2 storage; 16-frame steps into two alternating sets of device targets.
`roofline`: k_raster, algorithmic bytes 4*frags.i + 8*frags.o per launch over its
            CUDA-event duration, against MEASURED_PEAKS.json's HBM copy bandwidth.
`cpu_baseline`: the CPU oracle (1 thread, like the single-threaded reference) on a bounded
            sample of the same frames.
`--impl reference`: the oracle port on all host cores (frames are independent), same metric.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import retrofire_b200 as rf  # noqa: E402
from retrofire_b200 import scenes  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="bunny", choices=["bunny", "crates", "sprites", "small_tris"])
    ap.add_argument("--sharding", default="frames", choices=["frames", "tiles"],
                    help="frames: every rank renders its own frame batch (weak scaling); tiles: sort-first row bands of ONE large frame "
                         "per step + NCCL all_gather of the finished bands (strong scaling, SURVEY 8e)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "peer-root", "nccl"],
                    help="--sharding tiles: 'peer' fuses the exchange into the rasteriser (colour stores replicated into the peers over "
                         "NVLink, csrc/rf_peer.cuh); 'nccl' rasterises first and all_gathers the finished bands")
    ap.add_argument("--frames", type=int, default=128,
                    help="frames per step (frame batch); BASELINE config 5(ii) is a 1,024-frame batch over 8 GPUs = 128 per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--kernel-only", action="store_true", help="skip the e2e and cpu_baseline legs (for ncu runs)")
    return ap.parse_args()


# ---- workload ------------------------------------------------------------------------------------
def make_workload(name: str, frames: int):
    """Returns (scene template, list of per-frame DrawCall lists)."""
    if name == "bunny":
        base = scenes.bunny(subdiv=2)
        per_frame = []
        for f in range(frames):
            th = 2.0 * math.pi * f / frames + 1.0
            sc = scenes.bunny(subdiv=2, theta=th)
            per_frame.append(sc.draws)
        desc = {"workload": "bunny_x16 79,488 tris Gouraud+depth 1920x1080 Xrgb8888, frame batch", "frames_per_step": frames,
                "resolution": [base.w, base.h]}
        return base, per_frame, desc
    if name == "crates":
        base = scenes.crates("1089")
        per_frame = [base.draws for _ in range(frames)]
        desc = {"workload": "crates 1,089 textured cubes + floor 3840x2160 Rgba8888, one draw per cube", "frames_per_step": frames,
                "resolution": [base.w, base.h]}
        return base, per_frame, desc
    if name == "small_tris":
        base = scenes.small_tris(1_000_000)
        per_frame = [base.draws for _ in range(frames)]
        desc = {"workload": "1,000,000 small triangles (circumradius 1-8 px) 7680x4320 Rgba8888+depth", "frames_per_step": frames,
                "resolution": [base.w, base.h]}
        return base, per_frame, desc
    base = scenes.sprites(10000)
    per_frame = [scenes.sprites(10000, theta=1.0 + 0.1 * f).draws for f in range(frames)]
    desc = {"workload": "sprites 10k discs 1920x1080", "frames_per_step": frames, "resolution": [base.w, base.h]}
    return base, per_frame, desc


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows if len(r) > 2 + i)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---- reference arm / cpu baseline -------------------------------------------------------------------
def oracle_frames(base, per_frame, idx, threads: int):
    """Render frames `idx` with the CPU oracle; returns (seconds, frags_i, prims_i, frames)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import rfo
    rfo.load()

    def one(f):
        tgt = rfo.HostTarget(base.w, base.h, base.fmt, base.has_depth)
        tgt.clear(base.ctx.color_clear, base.ctx.depth_clear)
        fi = pi = 0
        for d in per_frame[f]:
            s = rfo.render(d, tgt)
            fi += s.frags.i
            pi += s.prims.i
        return fi, pi

    t0 = time.perf_counter()
    if threads <= 1:
        res = [one(f) for f in idx]
    else:
        with ThreadPoolExecutor(threads) as ex:
            res = list(ex.map(one, idx))
    dt = time.perf_counter() - t0
    return dt, sum(r[0] for r in res), sum(r[1] for r in res), len(idx)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, per_frame, desc = make_workload(args.workload, args.frames)
    cores = os.cpu_count() or 1
    # bounded sample per step: at most `cores` frames x 1 so a step stays around a second
    sample = list(range(min(args.frames, max(cores, 8))))
    for _ in range(max(args.warmup, 1)):
        oracle_frames(base, per_frame, sample[: max(1, len(sample) // 4)], cores)
    tot_t = tot_f = tot_p = tot_n = 0
    for _ in range(args.steps):
        dt, fi, pi, n = oracle_frames(base, per_frame, sample, cores)
        tot_t += dt; tot_f += fi; tot_p += pi; tot_n += n
    val = tot_f / tot_t / 1e6
    line = {
        "impl": "reference", "metric": "Mfragments/s", "value": val, "unit": "Mfragments/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": desc,
        "frames_per_s": tot_n / tot_t, "Mtriangles_per_s": tot_p / tot_t / 1e6,
        "cpu_baseline": {"value": val, "unit": "Mfragments/s", "cores": cores, "kind": "port",
                         "sample": f"{len(sample)} of {args.frames} frames per step, frames spread over {cores} host threads (oracle port; the Rust reference cannot be built here)"},
        "e2e": {"value": val, "unit": "Mfragments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---- B200 arm ------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream(device=local)

    base, per_frame, desc = make_workload(args.workload, args.frames)
    F = args.frames
    dev = rf.Device(local, stream=stream.cuda_stream)
    targets = [dev.framebuf(base.w, base.h, base.fmt, base.has_depth) for _ in range(F)]
    # resident geometry: one rf_mesh per distinct (prims, verts) pair
    mesh_cache = {}

    def resident(d: rf.DrawCall) -> rf.DrawCall:
        key = (d.prims.ctypes.data, d.verts.ctypes.data)
        if key not in mesh_cache:
            mesh_cache[key] = dev.mesh(d.prims, d.verts)
        import dataclasses
        return dataclasses.replace(d, mesh=mesh_cache[key])

    res_frames = [[resident(d) for d in draws] for draws in per_frame]
    single = all(len(dr) == 1 for dr in res_frames)
    uniforms = np.stack([dr[0].uniform for dr in res_frames]) if single else None

    def step_resident():
        for t in targets:
            t.clear(base.ctx)
        if single:
            dev.render_frames(res_frames[0][0], targets, uniforms)
        else:
            for t, draws in zip(targets, res_frames):
                dev.render_many(draws, t)    # the frame's render() calls in one crossing of the C ABI
        dev.flush()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also grows the arenas so the timed region never replays a pass)
    for _ in range(max(args.warmup, 3)):
        step_resident()
        dev.sync()
    dev.stats(reset=True)
    dev.profile(1)   # CUDA events around k_raster only: the timed passes keep their normal stream overlap
    torch.cuda.cudart().cudaProfilerStart()  # ncu --profile-from-start off: capture only the timed region

    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step_resident()
    ev1.record(stream)
    dev.sync()
    barrier()
    torch.cuda.cudart().cudaProfilerStop()
    ms = ev0.elapsed_time(ev1)
    sampler.stop_flag.set()
    st = dev.stats(reset=True)
    ktimes = dev.kernel_times()
    # kernel time shares: a few extra (untimed) steps with every kernel serialised and bracketed by events
    dev.profile(2)
    for _ in range(3):
        step_resident()
    kall = dev.kernel_times()
    dev.profile(0)
    _, launches_per_pass = dev.last_pass()

    # ---- single-frame latency, reported separately: clear + draw + wait, one frame per pass
    lat_ms = None
    if not args.kernel_only:
        one = res_frames[0]
        ts = []
        for k in range(30):
            dev.sync()
            t0 = time.perf_counter()
            targets[0].clear(base.ctx)
            dev.render_many(one, targets[0])
            dev.sync()
            ts.append(time.perf_counter() - t0)
        lat_ms = 1e3 * sorted(ts[5:])[len(ts[5:]) // 2]

    # ---- end to end: host geometry in (rf_render with host pointers: staged through pinned memory and
    # copied H2D inside the timed region), colour buffer of every frame out (D2H into page-locked Buf2 storage)
    Fe = min(F // 2, 16) if F >= 2 else 1   # frames per end-to-end step; two sets of device targets alternate (a swap chain)
    dt, shape = (np.uint32, (base.h, base.w)) if base.fmt == rf.FMT_XRGB8888 else (np.uint8, (base.h, base.w, 4))
    host_color = [[dev.pinned_empty(shape, dt) for _ in range(Fe)] for _ in range(2)]  # double-buffered Buf2 storage

    # the caller's vertex / index arrays live in page-locked memory (rf_host_alloc), as the bench contract asks
    import dataclasses as _dc
    pin_cache = {}

    def pinned_copy(a):
        key = a.ctypes.data
        if key not in pin_cache:
            b = dev.pinned_empty(a.shape, a.dtype)
            b[...] = a
            pin_cache[key] = b
        return pin_cache[key]

    e2e_frames = [[_dc.replace(d, prims=pinned_copy(d.prims), verts=pinned_copy(d.verts)) for d in per_frame[f]] for f in range(Fe)]

    def step_e2e(k=0):
        """One step through the reference-facing calls: clear + render() with host geometry for every frame,
        then the colour buffer of every frame is read back. Downloads run on the library's copy stream and
        overlap the next step's rendering; dev.sync() at the end of the timed region waits for all of them."""
        tg = targets[(k & 1) * Fe: (k & 1) * Fe + Fe] if F >= 2 * Fe else targets[:Fe]
        for f in range(Fe):
            tg[f].clear(base.ctx)
            dev.render_many(e2e_frames[f], tg[f])
        for f in range(Fe):
            tg[f].download_color_async(host_color[k & 1][f])

    e_steps = 0 if args.kernel_only else max(2, min(args.steps, 20))
    for k in range(2 if e_steps else 0):
        step_e2e(k)
    dev.sync()
    # wall-clock (host work is part of the end-to-end path): median of three repetitions of e_steps steps
    e_runs = []
    for rep in range(3 if e_steps else 1):
        dev.stats(reset=True)
        barrier()
        t0 = time.perf_counter()
        for k in range(e_steps):
            step_e2e(k)
        dev.sync()   # every queued pass and every download has completed: the pixels are in host memory
        barrier()
        e_runs.append(max(time.perf_counter() - t0, 1e-9))
    e_dt = sorted(e_runs)[len(e_runs) // 2]
    e_st = dev.stats(reset=True)
    h2d = sum(d.verts.nbytes + d.prims.nbytes for f in range(Fe) for d in per_frame[f])
    d2h = Fe * base.w * base.h * 4

    # ---- second metric configuration (BASELINE "crates 4K"): a short resident-geometry run, reported under "crates_4k"
    crates_info = None
    if args.workload == "bunny" and not args.kernel_only:
        cb, cpf, _ = make_workload("crates", 2)
        ct = [dev.framebuf(cb.w, cb.h, cb.fmt, cb.has_depth) for _ in range(2)]
        cres = [[resident(d) for d in draws] for draws in cpf]

        def step_crates():
            for t, draws in zip(ct, cres):
                t.clear(cb.ctx)
                dev.render_many(draws, t)
            dev.flush()

        for _ in range(3):
            step_crates()
            dev.sync()
        dev.stats(reset=True)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(stream)
        csteps = 10
        for _ in range(csteps):
            step_crates()
        c1.record(stream)
        dev.sync()
        cms = c0.elapsed_time(c1)
        cst = dev.stats(reset=True)
        crates_info = {"workload": "crates 1,089 cubes + floor 3840x2160 Rgba8888, one draw per visible cube", "draws_per_frame": len(cres[0]),
                       "frames_per_s": 2 * csteps / (cms * 1e-3), "Mfragments_per_s": cst.frags.i / (cms * 1e-3) / 1e6,
                       "ms_per_frame": cms / (2 * csteps), "frags_i_per_frame": cst.frags.i // (2 * csteps), "n_gpus": 1}

    # ---- reduce over ranks
    t_ms, e_s = ms, e_dt
    frags_i, frags_o, prims_i = st.frags.i, st.frags.o, st.prims.i
    e_frags = e_st.frags.i
    if world > 1:
        tt = torch.tensor([ms, e_dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_ms, e_s = float(tt[0]), float(tt[1])
        cc = torch.tensor([frags_i, frags_o, prims_i, e_frags], device="cuda", dtype=torch.int64)
        dist.all_reduce(cc, op=dist.ReduceOp.SUM)
        frags_i, frags_o, prims_i, e_frags = (int(x) for x in cc)

    if rank == 0:
        peak, peak_src = peaks()
        r_ns, r_n = ktimes["k_raster"]
        # rank-0 kernel: algorithmic bytes of ITS launches
        alg_bytes_launch = (4 * st.frags.i + 8 * st.frags.o) / max(r_n, 1)
        r_avg_s = r_ns * 1e-9 / max(r_n, 1)
        achieved = alg_bytes_launch / r_avg_s / 1e9 if r_avg_s > 0 else 0.0
        kshare = {k: round(v[0] / max(sum(x[0] for x in kall.values()), 1), 4) for k, v in kall.items()}
        # whole-frame algorithmic bytes (SURVEY §8d B_alg) over the whole step time, for context
        geom = base.geometry_bytes() if single else sum(d.verts.shape[0] * 4 * (3 + d.shader.lanes) + d.prims.shape[0] * 12 for d in per_frame[0])
        b_alg_step = F * (geom + 8 * base.w * base.h) + (4 * st.frags.i + 8 * st.frags.o) / args.steps
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r01_raster_traffic.json")
        if os.path.exists(tp):  # dram bytes of one k_raster launch from the committed ncu --set full capture of this command
            tj = json.load(open(tp))
            if tj.get("workload") == args.workload and tj.get("frames_per_pass") == F:
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        line = {
            "metric": "Mfragments/s", "value": frags_i / (t_ms * 1e-3) / 1e6, "unit": "Mfragments/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": t_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(desc, l2="inputs larger than L2: %d targets x %.1f MB colour+depth per step" % (F, base.w * base.h * 8 / 1e6),
                           parallelism=f"frame-sharded x{world}"),
            "frames_per_s": world * F * args.steps / (t_ms * 1e-3), "Mtriangles_per_s": prims_i / (t_ms * 1e-3) / 1e6,
            "frags_i_per_step": frags_i // args.steps, "frags_o_per_step": frags_o // args.steps,
            "e2e": {"value": e_frags / e_s / 1e6, "unit": "Mfragments/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "frames_per_step": Fe, "frames_per_s": world * Fe * e_steps / e_s},
            "gpu_launches": int(args.steps * (launches_per_pass)),
            "roofline": {"bound": "hbm", "kernel": "k_raster", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel_ms_avg": r_avg_s * 1e3,
                         "alg_bytes_per_launch": alg_bytes_launch, "kernel_time_share": kshare,
                         "pass_alg_GBps": b_alg_step / (t_ms * 1e-3 / args.steps) / 1e9},
            "clocks": sampler.summary(),
        }
        if crates_info is not None:
            line["crates_4k"] = crates_info
        if lat_ms is not None:
            line["single_frame_latency_ms"] = lat_ms  # wall clock of clear + render + sync for ONE frame (not batched)
        if world == 1 and not args.kernel_only:
            # CPU baseline: oracle, 1 thread (the reference is single-threaded), bounded sample
            n = 0
            t_used = fi_tot = 0.0
            while t_used < args.cpu_seconds and n < 100000:
                dt, fi, _, _ = oracle_frames(base, per_frame, [n % F], 1)
                t_used += dt; fi_tot += fi; n += 1
            line["cpu_baseline"] = {"value": fi_tot / t_used / 1e6, "unit": "Mfragments/s", "cores": 1, "kind": "port",
                                    "sample": f"{n} frames (cycling the step's {F}), CPU oracle single-threaded, {t_used:.1f} s",
                                    "frames_per_s": n / t_used}
        print(json.dumps(line), flush=True)
    dev.close()
    if world > 1:
        dist.destroy_process_group()


def run_tiles(args):
    """Sort-first: geometry replicated, rank r rasterises its row band of one large frame, bands gathered with NCCL."""
    import torch
    import torch.distributed as dist
    from retrofire_b200 import shard
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream(device=local)
    base, per_frame, desc = make_workload(args.workload, 1)
    dev = rf.Device(local, stream=stream.cuda_stream)
    bands = shard.row_bands(base.h, world)
    dev.set_row_band(*bands[rank])
    fb = dev.framebuf(base.w, base.h, base.fmt, base.has_depth)
    import dataclasses
    draws = [dataclasses.replace(d, mesh=dev.mesh(d.prims, d.verts)) for d in per_frame[0]]
    color = shard.target_tensor(fb)
    peer = world > 1 and args.exchange.startswith("peer")
    if peer:   # "peer": every rank ends with the whole frame; "peer-root": only rank 0 collects it
        shard.attach_peers(dev, [fb], rank, world, root=0 if args.exchange == "peer-root" else None)

    def step():
        fb.clear(base.ctx)
        for d in draws:
            dev.render(d, fb)
        dev.flush()
        if world > 1 and not peer:
            with torch.cuda.stream(stream):
                shard.gather_bands(color, bands, rank)

    for _ in range(max(args.warmup, 3)):
        step()
        dev.sync()
    dev.stats(reset=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    dev.sync()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    st = dev.stats(reset=True)
    fi, pi = st.frags.i, st.prims.i
    # where the step goes (untimed extra steps): every kernel serialised and bracketed by events; the gather alone
    dev.profile(2)
    for _ in range(3):
        step()
    dev.sync()
    kall = dev.kernel_times()
    dev.profile(0)
    kernel_ms = {k: round(v[0] * 1e-6 / max(v[1], 1), 4) for k, v in kall.items()}
    gather_ms = 0.0
    if world > 1:
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            g0.record(stream)
            for _ in range(5):
                shard.gather_bands(color, bands, rank)
            g1.record(stream)
        torch.cuda.synchronize()
        gather_ms = g0.elapsed_time(g1) / 5
    if world > 1:
        tt = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt[0])
        cc = torch.tensor([fi], device="cuda", dtype=torch.int64)
        dist.all_reduce(cc, op=dist.ReduceOp.SUM)
        fi = int(cc[0])
    if rank == 0:
        print(json.dumps({
            "metric": "Mfragments/s", "value": fi / (ms * 1e-3) / 1e6, "unit": "Mfragments/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict(desc, parallelism=f"sort-first row bands x{world}, " + ("colour stores replicated into the peers over NVLink inside k_raster"
                                                                                   if peer else "NCCL all_gather of finished bands"), bands=bands,
                           l2="inputs larger than L2: 265 MB colour+depth target + 84 MB geometry per step"),
            "frames_per_s": args.steps / (ms * 1e-3), "Mtriangles_per_s": pi / (ms * 1e-3) / 1e6,
            "exchange": "none" if world == 1 else args.exchange,
            "gather_bytes_per_step": 0 if world == 1 or peer else base.w * base.h * 4 * (world - 1) // world,
            "nccl_gather_ms_alone": gather_ms, "kernel_ms_rank0": kernel_ms}), flush=True)
    dev.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.sharding == "tiles":
        run_tiles(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
