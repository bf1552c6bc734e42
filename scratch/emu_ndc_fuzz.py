"""Ad hoc fuzz under the SIMT emulation (no GPU): NDC triangle soups up to the size of the target, random target sizes (1..5000 px, extreme
aspects included), pixel formats and Context flags; device path vs oracle, error statuses included. Usage: python scratch/emu_ndc_fuzz.py <seed> <cases>"""
import ctypes as C, os, sys, time, dataclasses
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests/emu"); os.chdir("/root/repo")
import numpy as np
import retrofire_b200 as rf
from retrofire_b200 import _ffi, scenes
from oracle import rfo
from tests import test_gpu_3_adversarial as G
from tests.parity import run_gpu, run_oracle, assert_parity
import build_emu
rfo.build(); rfo.load()
lib = C.CDLL(build_emu.build())
for name, (res, args) in _ffi.SYMBOLS.items():
    fn = getattr(lib, name); fn.restype = res; fn.argtypes = args
saved, _ffi._lib = _ffi._lib, lib
dev = rf.Device(0)
_ffi._lib = saved
g = np.random.default_rng(int(sys.argv[1]))
bad = 0
fmts = [rf.FMT_RGBA8888, rf.FMT_XRGB8888, rf.FMT_RGB888, rf.FMT_RGB565, rf.FMT_RGBA4444, rf.FMT_ARGB8888, rf.FMT_BGRA8888]
for it in range(int(sys.argv[2])):
    w, h = int(g.integers(1, 900)), int(g.integers(1, 500))
    if g.integers(0, 4) == 0: w, h = (int(g.integers(1, 5000)), int(g.integers(1, 40))) if g.integers(0, 2) else (int(g.integers(1, 40)), int(g.integers(1, 5000)))
    persp = bool(g.integers(0, 2))
    sc = G.ndc_soup(int(g.integers(1, 700)), w, h, seed=int(g.integers(1, 1 << 30)), persp=persp, extent=float(g.choice([0.05, 0.3, 1.0, 2.5])))
    d = sc.draws[0]
    ctx = rf.Context(face_cull=[None, rf.FaceCull.Back, rf.FaceCull.Front][g.integers(0, 3)], depth_test=[None, rf.Ordering.Less, rf.Ordering.Greater, rf.Ordering.Equal][g.integers(0, 4)],
                     depth_write=bool(g.integers(0, 4)), color_write=bool(g.integers(0, 6)), depth_sort=[None, None, rf.DepthSort.BackToFront, rf.DepthSort.FrontToBack][g.integers(0, 4)])
    if ctx.depth_test == rf.Ordering.Greater: ctx.depth_clear = 0.001
    d = dataclasses.replace(d, face_cull=ctx.face_cull or 0, depth_test=ctx.depth_test or 0, color_write=ctx.color_write, depth_write=ctx.depth_write, depth_sort=0 if ctx.depth_sort is None else int(ctx.depth_sort))
    sc = scenes.Scene(sc.name, w, h, fmts[g.integers(0, len(fmts))], bool(g.integers(0, 5) > 0), ctx, [d])
    try:
        want = run_oracle(rfo, sc)
    except rf.RetrofireError as e:
        try:
            run_gpu(dev, sc); print(it, sc.name, "oracle error but device ok:", e); bad += 1
        except rf.RetrofireError as ge:
            if ge.status != e.status: print(it, "status differs", ge, e); bad += 1
        continue
    try:
        assert_parity(run_gpu(dev, sc), want, name=sc.name)
    except Exception as e:
        print(it, sc.name, w, h, persp, ctx, "FAIL", str(e)[:300]); bad += 1
print("done", sys.argv[1:], "bad", bad)
