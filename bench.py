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
            page-locked Buf2 storage; steps of F/2 frames into two alternating halves of the device targets; `frac_of_pcie` = its D2H
            rate over the device-to-host ceiling measured in the same run with every rank copying at once.
`roofline`: k_raster, algorithmic bytes 4*frags.i + 8*frags.o per launch over its
            CUDA-event duration, against MEASURED_PEAKS.json's HBM copy bandwidth.
`cpu_baseline`: the CPU oracle (1 thread, like the single-threaded reference) on a bounded
            sample of the same frames.
`crates_4k`: the second configuration BASELINE.json's metric names (crates 3840x2160, 32-frame batches), same objects.
`sort_first_8k` (N > 1): one 8K frame of 1 M small triangles, sort-first over the ranks, both exchange forms.
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
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks line). NVML in-process (a sample
    costs microseconds, so even a 50 ms region of 20 steps gets dozens); `nvidia-smi` as a subprocess only when NVML cannot be loaded
    (a subprocess takes longer than a short timed region: round 1's 8-rank line had no sample at all)."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int, uuid: str | None = None):
        super().__init__(daemon=True)
        self.index, self.uuid, self.stop_flag, self.rows = index, uuid, threading.Event(), []   # rows: (sm_mhz, max_mhz, [reason flags])
        self.source = None

    def _nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        h = None
        if self.uuid:
            try:
                h = nv.nvmlDeviceGetHandleByUUID(self.uuid if self.uuid.startswith("GPU-") else "GPU-" + self.uuid)
            except Exception:
                h = None
        if h is None:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [v for v in vis.split(",") if v.strip().isdigit()]
            h = nv.nvmlDeviceGetHandleByIndex(int(ids[self.index]) if self.index < len(ids) else self.index)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        masks = [nv.nvmlClocksEventReasonHwSlowdown, nv.nvmlClocksEventReasonHwThermalSlowdown,
                 nv.nvmlClocksEventReasonSwThermalSlowdown, nv.nvmlClocksEventReasonSwPowerCap]

        def sample():
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            return nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), mx, [bool(r & m) for m in masks]
        sample()
        return sample

    def _smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

        def sample():
            out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=5).stdout.strip()
            r = [s.strip() for s in out.split(",")]
            return int(r[0]), int(r[1]), [x.lower().startswith("active") for x in r[2:6]]
        return sample

    def prepare(self):
        """Open NVML before the timed region starts (library load is not part of the sampling)."""
        try:
            self.sample, self.source = self._nvml(), "nvml"
        except Exception:
            self.sample, self.source = self._smi(), "nvidia-smi"
        return self

    def run(self):
        period = 0.002 if self.source == "nvml" else 0.1
        while True:
            try:
                self.rows.append(self.sample())
            except Exception:
                pass
            if self.stop_flag.wait(period):   # the last sample is taken before the stop is seen: at least one lies in the region
                break

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": self.source}
        sm = sorted(r[0] for r in self.rows)
        reasons = [n for i, n in enumerate(self.NAMES) if any(r[2][i] for r in self.rows if len(r[2]) > i)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.rows[0][1], "reasons": reasons, "samples": len(self.rows),
                "source": self.source}


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
        "frames_sampled": len(sample), "threads": cores,
        "cpu_baseline": {"value": val, "unit": "Mfragments/s", "cores": cores, "kind": "port",
                         "sample": f"{len(sample)} of {args.frames} frames per step, frames spread over {cores} host threads (oracle port; the Rust reference cannot be built here)"},
        "e2e": {"value": val, "unit": "Mfragments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---- B200 arm ------------------------------------------------------------------------------------------
def csrc_sha() -> str:
    """Hash of the kernel sources: a committed ncu traffic figure is only printed for the tree it was captured from."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "retrofire_b200", "csrc")
    for name in sorted(os.listdir(d)):
        with open(os.path.join(d, name), "rb") as f:
            h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()[:16]


def committed_traffic(workload: str, frames: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE k_raster launch of this command, from the committed capture
    (profiles/r02_raster_traffic.json, written by profiles/make_traffic.py from an ncu run) — or None when the capture is of
    another tree (csrc hash), workload or batch size: a stale figure is never printed."""
    tp = os.path.join(ROOT, "profiles", "r02_raster_traffic.json")
    if not os.path.exists(tp):
        return None, None
    tj = json.load(open(tp))
    e = tj.get("captures", {}).get(f"{workload}:{frames}")
    if not e or tj.get("csrc_sha") != csrc_sha():
        return None, None
    return int(e["dram_bytes_read"] + e["dram_bytes_write"]), e.get("pass_dram_bytes")


def measure_pcie(torch, dist, world, seconds: float = 1.0):
    """Device-to-host DMA ceiling of this box with all `world` ranks copying at once (page-locked destination), GB/s per rank
    and aggregate: the denominator of the end-to-end leg, whose every frame ends with one colour buffer crossing PCIe.
    Copies of 64 MB are timed one by one with CUDA events for `seconds`; the first third is warm-up (the link leaves its idle
    state under load — a ceiling read during the ramp came out BELOW the rate the end-to-end leg then reached), the figure is
    the median rate of the rest."""
    n = 64 << 20
    src = torch.empty(n, dtype=torch.uint8, device="cuda")
    dst = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    rates = []
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(9)]
        evs[0].record()
        for k in range(8):
            dst.copy_(src, non_blocking=True)
            evs[k + 1].record()
        torch.cuda.synchronize()
        now = time.perf_counter() - t0
        rates += [(now, n / (evs[k].elapsed_time(evs[k + 1]) * 1e-3) / 1e9) for k in range(8)]
    late = sorted(r for t, r in rates if t >= seconds / 3) or sorted(r for t, r in rates)
    mine = late[len(late) // 2]
    tot = mine
    if world > 1:
        t = torch.tensor([mine], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        tot = float(t[0])
    return mine, tot


def measure(ctx, workload: str, F: int, steps: int, warmup: int, e2e: bool, cpu_seconds: float, latency: bool):
    """One workload through the device path. Returns the fields of a bench line for it (rank-local; reduce() folds ranks)."""
    import dataclasses
    torch, dev, stream, barrier = ctx["torch"], ctx["dev"], ctx["stream"], ctx["barrier"]
    base, per_frame, desc = make_workload(workload, F)
    targets = [dev.framebuf(base.w, base.h, base.fmt, base.has_depth) for _ in range(F)]
    mesh_cache = ctx["mesh_cache"]

    def resident(d: rf.DrawCall) -> rf.DrawCall:  # resident geometry: one rf_mesh per distinct (prims, verts) pair
        key = (d.prims.ctypes.data, d.verts.ctypes.data)
        if key not in mesh_cache:
            mesh_cache[key] = dev.mesh(d.prims, d.verts)
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

    # ---- warm-up (also grows the arenas so the timed region never replays a pass)
    for _ in range(max(warmup, 3)):
        step_resident()
        dev.sync()
    dev.stats(reset=True)
    dev.profile(1)   # CUDA events around k_raster only: the timed passes keep their normal stream overlap
    if ctx["primary"]:
        torch.cuda.cudart().cudaProfilerStart()  # ncu --profile-from-start off: capture only the timed region
    try:
        uuid = str(torch.cuda.get_device_properties(torch.cuda.current_device()).uuid)
    except Exception:
        uuid = None
    sampler = ClockSampler(ctx["local"], uuid).prepare()
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        step_resident()
    ev1.record(stream)
    dev.sync()
    barrier()
    if ctx["primary"]:
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
    out = {"base": base, "desc": desc, "F": F, "steps": steps, "ms": ms, "stats": st, "ktimes": ktimes, "kall": kall,
           "launches_per_pass": launches_per_pass, "clocks": sampler.summary(), "single": single, "per_frame": per_frame}

    # ---- single-frame latency, reported separately: clear + draw + wait, one frame per pass
    if latency:
        one = res_frames[0]
        ts = []
        for k in range(30):
            dev.sync()
            t0 = time.perf_counter()
            targets[0].clear(base.ctx)
            dev.render_many(one, targets[0])
            dev.sync()
            ts.append(time.perf_counter() - t0)
        out["latency_ms"] = 1e3 * sorted(ts[5:])[len(ts[5:]) // 2]

    # ---- end to end: host geometry in (rf_render with HOST pointers to page-locked arrays: DMA'd inside the timed region),
    # colour buffer of every frame out (D2H into page-locked Buf2 storage). Two halves of the step's targets alternate (a swap
    # chain), so that the downloads of one half overlap the rendering of the other.
    if e2e:
        Fe = max(1, F // 2)
        dt, shape = (np.uint32, (base.h, base.w)) if base.fmt == rf.FMT_XRGB8888 else (np.uint8, (base.h, base.w, 4))
        host_color = [[dev.pinned_empty(shape, dt) for _ in range(Fe)] for _ in range(2)]  # double-buffered Buf2 storage
        pin_cache = {}

        def pinned_copy(a):
            key = a.ctypes.data
            if key not in pin_cache:
                b = dev.pinned_empty(a.shape, a.dtype)
                b[...] = a
                pin_cache[key] = b
            return pin_cache[key]

        e2e_frames = [[dataclasses.replace(d, prims=pinned_copy(d.prims), verts=pinned_copy(d.verts)) for d in per_frame[f]] for f in range(Fe)]

        def step_e2e(k=0):
            tg = targets[(k & 1) * Fe: (k & 1) * Fe + Fe] if F >= 2 * Fe else targets[:Fe]
            for f in range(Fe):
                tg[f].clear(base.ctx)
                dev.render_many(e2e_frames[f], tg[f])
            for f in range(Fe):
                tg[f].download_color_async(host_color[k & 1][f])

        e_steps = max(4, min(2 * steps, 12))
        for k in range(2):
            step_e2e(k)
        dev.sync()
        e_runs = []   # wall clock (host work is part of the end-to-end path): median of three repetitions
        for rep in range(3):
            dev.stats(reset=True)
            barrier()
            t0 = time.perf_counter()
            for k in range(e_steps):
                step_e2e(k)
            dev.sync()   # every queued pass and every download has completed: the pixels are in host memory
            barrier()
            e_runs.append(max(time.perf_counter() - t0, 1e-9))
        out["e2e"] = {"seconds": sorted(e_runs)[1], "stats": dev.stats(reset=True), "steps": e_steps, "Fe": Fe,
                      "h2d": sum(d.verts.nbytes + d.prims.nbytes for f in range(Fe) for d in per_frame[f]), "d2h": Fe * base.w * base.h * 4}

    if cpu_seconds > 0:  # CPU baseline: oracle, 1 thread (the reference is single-threaded), bounded sample
        n = 0
        t_used = fi_tot = 0.0
        while t_used < cpu_seconds and n < 100000:
            dt_, fi, _, _ = oracle_frames(base, per_frame, [n % F], 1)
            t_used += dt_; fi_tot += fi; n += 1
        out["cpu"] = {"value": fi_tot / t_used / 1e6, "unit": "Mfragments/s", "cores": 1, "kind": "port",
                      "sample": f"{n} frames (cycling the step's {F}), CPU oracle single-threaded, {t_used:.1f} s", "frames_per_s": n / t_used}
    for t in targets:
        t._destroy()
        dev._targets.remove(t)
    return out


def fold(ctx, m, workload: str, pcie):
    """Fold a measure() result over the ranks (max of times, sum of counters) into bench-line fields (valid on rank 0)."""
    torch, dist, world = ctx["torch"], ctx["dist"], ctx["world"]
    base, F, steps, st = m["base"], m["F"], m["steps"], m["stats"]
    t_ms = m["ms"]
    frags_i, frags_o, prims_i = st.frags.i, st.frags.o, st.prims.i
    e = m.get("e2e")
    e_s, e_frags = (e["seconds"], e["stats"].frags.i) if e else (0.0, 0)
    if world > 1:
        tt = torch.tensor([t_ms, e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_ms, e_s = float(tt[0]), float(tt[1])
        cc = torch.tensor([frags_i, frags_o, prims_i, e_frags], device="cuda", dtype=torch.int64)
        dist.all_reduce(cc, op=dist.ReduceOp.SUM)
        frags_i, frags_o, prims_i, e_frags = (int(x) for x in cc)
    peak, peak_src = peaks()
    r_ns, r_n = m["ktimes"]["k_raster"]
    alg_bytes_launch = (4 * st.frags.i + 8 * st.frags.o) / max(r_n, 1)   # rank-0 kernel: algorithmic bytes of ITS launches
    r_avg_s = r_ns * 1e-9 / max(r_n, 1)
    achieved = alg_bytes_launch / r_avg_s / 1e9 if r_avg_s > 0 else 0.0
    kall = m["kall"]
    kshare = {k: round(v[0] / max(sum(x[0] for x in kall.values()), 1), 4) for k, v in kall.items()}
    # whole-frame algorithmic bytes (SURVEY §8d B_alg: geometry in + clear + 4 B per input fragment + 8 B per written one) of
    # one rank's step over that step's time: BASELINE.md's "% of HBM peak"
    geom = sum(d.verts.shape[0] * 4 * (3 + d.shader.lanes) + d.prims.shape[0] * 12 for d in m["per_frame"][0])
    b_alg_step = F * (geom + (8 if base.has_depth else 4) * base.w * base.h) + (4 * st.frags.i + 8 * st.frags.o) / steps
    pass_gbps = b_alg_step / (t_ms * 1e-3 / steps) / 1e9
    traffic, pass_traffic = committed_traffic(workload, F)
    f = {
        "value": frags_i / (t_ms * 1e-3) / 1e6, "ms_per_step": t_ms / steps,
        "config": dict(m["desc"], l2="inputs larger than L2: %d targets x %.1f MB colour+depth per step" % (F, base.w * base.h * 8 / 1e6),
                       parallelism=f"frame-sharded x{world}"),
        "frames_per_s": world * F * steps / (t_ms * 1e-3), "Mtriangles_per_s": prims_i / (t_ms * 1e-3) / 1e6,
        "frags_i_per_step": frags_i // steps, "frags_o_per_step": frags_o // steps,
        "gpu_launches": int(steps * m["launches_per_pass"]),
        "roofline": {"bound": "hbm", "kernel": "k_raster", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": "profiles/r02_raster_traffic.json (ncu capture of this tree, csrc hash checked)" if traffic else None,
                     "peak_source": peak_src, "kernel_ms_avg": r_avg_s * 1e3, "alg_bytes_per_launch": alg_bytes_launch,
                     "kernel_time_share": kshare,
                     "pass_alg_GBps": pass_gbps, "frac_pass": pass_gbps / peak, "pass_alg_bytes": b_alg_step, "pass_traffic": pass_traffic},
        "clocks": m["clocks"],
    }
    if e:
        frames_s = world * e["Fe"] * e["steps"] / e_s
        f["e2e"] = {"value": e_frags / e_s / 1e6, "unit": "Mfragments/s", "h2d_bytes_per_step": int(e["h2d"]), "d2h_bytes_per_step": int(e["d2h"]),
                    "frames_per_step": e["Fe"], "frames_per_s": frames_s}
        if pcie:
            d2h_gbps = frames_s * base.w * base.h * 4 / 1e9
            f["e2e"].update({"d2h_GBps": d2h_gbps, "pcie_d2h_GBps_all_ranks": pcie[1], "frac_of_pcie": d2h_gbps / pcie[1],
                             "bound": "host link: every frame ends with its colour buffer crossing PCIe"})
    if "latency_ms" in m:
        f["single_frame_latency_ms"] = m["latency_ms"]  # wall clock of clear + render + sync for ONE frame (not batched)
    if "cpu" in m:
        f["cpu_baseline"] = m["cpu"]
    return f


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
    dev = rf.Device(local, stream=stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = {"torch": torch, "dist": dist, "dev": dev, "stream": stream, "barrier": barrier, "world": world, "rank": rank, "local": local,
           "mesh_cache": {}, "primary": True}
    full = not args.kernel_only
    pcie = measure_pcie(torch, dist, world) if full else None
    m = measure(ctx, args.workload, args.frames, args.steps, args.warmup, e2e=full, cpu_seconds=args.cpu_seconds if full and world == 1 else 0.0, latency=full)
    f = fold(ctx, m, args.workload, pcie)
    line = {"metric": "Mfragments/s", "value": f.pop("value"), "unit": "Mfragments/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": f.pop("ms_per_step"), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic"}
    line.update(f)
    # ---- the second configuration BASELINE.json's metric names, "crates 4K", measured the same way (value, roofline, e2e,
    # cpu_baseline) on a 32-frame batch (8-frame batches leave the pass pipeline half empty: 4,580 vs 5,250 frames/s)
    if args.workload == "bunny" and full:
        ctx["primary"] = False
        mc = measure(ctx, "crates", 32, 20, 3, e2e=True, cpu_seconds=min(args.cpu_seconds, 4.0) if world == 1 else 0.0, latency=True)
        fc = fold(ctx, mc, "crates", pcie)
        fc["unit"] = "Mfragments/s"
        fc["draws_per_frame"] = len(mc["per_frame"][0])
        line["crates_4k"] = fc
    dev.close()
    # ---- N > 1: the other way the path shards (north_star config 5-i): ONE 8K frame of 1 M small triangles,
    # sort-first over the ranks, so that the scaling record carries the mode that has an exchange step
    if world > 1 and full and args.workload == "bunny":
        a2 = argparse.Namespace(**vars(args))
        a2.workload, a2.steps, a2.warmup = "small_tris", 20, 3
        sf = {}
        for ex in ("peer", "nccl"):
            a2.exchange = ex
            r = run_tiles(a2, emit=False, init=False)
            if rank == 0 and r:
                sf[ex] = {k: r[k] for k in ("value", "ms_per_step", "frames_per_s", "exchange", "gather_bytes_per_step", "nccl_gather_ms_alone", "kernel_ms_rank0")}
                sf["workload"] = r["config"]["workload"]
        if rank == 0:
            line["sort_first_8k"] = sf
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_tiles(args, emit: bool = True, init: bool = True):
    """Sort-first: geometry replicated, rank r rasterises its row band of one large frame, bands gathered with NCCL."""
    import torch
    import torch.distributed as dist
    from retrofire_b200 import shard
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1 and init:
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
    result = None
    if rank == 0:
        result = ({
            "metric": "Mfragments/s", "value": fi / (ms * 1e-3) / 1e6, "unit": "Mfragments/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict(desc, parallelism=f"sort-first row bands x{world}, " + ("colour stores replicated into the peers over NVLink inside k_raster"
                                                                                   if peer else "NCCL all_gather of finished bands"), bands=bands,
                           l2="inputs larger than L2: 265 MB colour+depth target + 84 MB geometry per step"),
            "frames_per_s": args.steps / (ms * 1e-3), "Mtriangles_per_s": pi / (ms * 1e-3) / 1e6,
            "exchange": "none" if world == 1 else args.exchange,
            "gather_bytes_per_step": 0 if world == 1 or peer else base.w * base.h * 4 * (world - 1) // world,
            "nccl_gather_ms_alone": gather_ms, "kernel_ms_rank0": kernel_ms})
        if emit:
            print(json.dumps(result), flush=True)
    dev.close()
    if world > 1 and init:
        dist.destroy_process_group()
    return result


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
