"""Shared helpers: run a Scene through the CUDA path (C ABI) and through the CPU oracle, compare."""
import numpy as np

import retrofire_b200 as rf


def run_oracle(oracle, sc, band=None):
    tgt = oracle.HostTarget(sc.w, sc.h, sc.fmt, sc.has_depth)
    if band is not None:
        tgt.band = band
    if sc.clear:
        tgt.clear(sc.ctx.color_clear, sc.ctx.depth_clear)
    stats = rf.Stats()
    for d in sc.draws:
        stats += oracle.render(d, tgt)
    return tgt.host_color(), tgt.depth, stats


def run_gpu(dev, sc, per_draw_sync=False):
    fb = dev.framebuf(sc.w, sc.h, sc.fmt, sc.has_depth)
    try:
        if sc.clear:
            fb.clear(sc.ctx)
        dev.stats(reset=True)
        if per_draw_sync:
            stats = rf.Stats()
            for d in sc.draws:
                stats += dev.render(d, fb, want_stats=True)
        else:
            for d in sc.draws:
                dev.render(d, fb)
            stats = dev.stats(reset=True)
        color = fb.download_color()
        depth = fb.download_depth() if sc.has_depth else None
    finally:
        fb._destroy()
        dev._targets.remove(fb)
    return color, depth, stats


def depth_equal(a, b):
    """Bit-equal depth buffers, NaNs equal to NaNs whatever their payload (the NaN contract)."""
    a = np.asarray(a, dtype=np.float32); b = np.asarray(b, dtype=np.float32)
    return a.shape == b.shape and bool((((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b)))).all())


def assert_parity(got, want, *, color_tol=0, name=""):
    gc, gd, gs = got
    wc, wd, ws = want
    if wd is not None:
        # NaN contract: a NaN depth is a NaN at the same pixel on both sides; its payload is not part of the result
        # (Rust leaves NaN bit patterns unspecified; x86 generates 0xFFC00000, CUDA 0x7FFFFFFF, propagation keeps an input's).
        # test_generated_nan_depth_has_the_host_bit_pattern pins the bits the device writes for NaNs it generates.
        db = (gd.view(np.uint32) != wd.view(np.uint32)) & ~(np.isnan(gd) & np.isnan(wd))
        assert not db.any(), f"{name}: {int(db.sum())} depth values differ (first at {np.argwhere(db)[0]})"
    if color_tol == 0:
        cb = gc != wc
        if cb.ndim == 3:
            cb = cb.any(axis=2)
        assert not cb.any(), f"{name}: {int(cb.sum())} pixels differ in colour (first at {np.argwhere(cb)[0]}: got {gc[tuple(np.argwhere(cb)[0])]}, want {wc[tuple(np.argwhere(cb)[0])]})"
    else:
        d = np.abs(gc.astype(np.int64) - wc.astype(np.int64))
        assert d.max() <= color_tol, f"{name}: colour differs by {d.max()} LSB"
    assert gs.counters() == ws.counters(), f"{name}: Stats differ: got {gs.counters()} want {ws.counters()}"
