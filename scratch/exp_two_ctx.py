import sys, time, math, dataclasses
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import retrofire_b200 as rf
from retrofire_b200 import scenes
F = 32
base = scenes.bunny(subdiv=2)
unis = np.stack([scenes.bunny(subdiv=2, theta=2*math.pi*f/F + 1.0).draws[0].uniform for f in range(F)])
def mk(n):
    dev = rf.Device(0)
    tg = [dev.framebuf(base.w, base.h, base.fmt, True) for _ in range(n)]
    d = base.draws[0]
    mesh = dev.mesh(d.prims, d.verts)
    call = dataclasses.replace(d, mesh=mesh)
    return dev, tg, call
def run(devs, steps):
    for dev, tg, call, u in devs:
        for _ in range(3):
            for t in tg: t.clear(base.ctx)
            dev.render_frames(call, tg, u); dev.flush(); dev.sync()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        for dev, tg, call, u in devs:
            for t in tg: t.clear(base.ctx)
            dev.render_frames(call, tg, u); dev.flush()
    for dev, *_ in devs: dev.sync()
    return (time.perf_counter() - t0) / steps * 1e3
for nctx in (1, 2, 4):
    n = F // nctx
    devs = []
    for k in range(nctx):
        dev, tg, call = mk(n)
        devs.append((dev, tg, call, unis[k*n:(k+1)*n]))
    print(nctx, "contexts x", n, "frames:", round(run(devs, 50), 3), "ms per 32 frames")
    for dev, *_ in devs: dev.close()
