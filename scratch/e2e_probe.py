import sys, time, math
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import retrofire_b200 as rf
from retrofire_b200 import scenes
F = 16
base = scenes.bunny(subdiv=2)
frames = [scenes.bunny(subdiv=2, theta=2*math.pi*f/32 + 1.0).draws for f in range(F)]
dev = rf.Device(0)
tgs = [[dev.framebuf(base.w, base.h, base.fmt, True) for _ in range(F)] for _ in range(2)]
tg = tgs[0]
host = [[dev.pinned_empty((base.h, base.w), np.uint32) for _ in range(F)] for _ in range(2)]
import dataclasses
def pin(x):
    y = dev.pinned_empty(x.shape, x.dtype); y[...] = x; return y
pv, pp = pin(frames[0][0].verts), pin(frames[0][0].prims)
if len(sys.argv) > 1:
    frames = [[dataclasses.replace(d, verts=pv, prims=pp) for d in fr] for fr in frames]
a = frames[0][0].verts; b = np.empty_like(a)
t0 = time.perf_counter()
for _ in range(50): np.copyto(b, a)
print("np.copyto %.2f GB/s" % (a.nbytes * 50 / (time.perf_counter() - t0) / 1e9))
t0 = time.perf_counter()
for _ in range(200): tg[0].clear(base.ctx)
print("clear call %.1f us" % ((time.perf_counter() - t0) / 200 * 1e6)); dev.sync()
def step(k, t):
    tg = tgs[k & 1]
    t0 = time.perf_counter()
    for f in range(F):
        tg[f].clear(base.ctx)
        for d in frames[f]: dev.render(d, tg[f])
    t1 = time.perf_counter()
    dev.flush()
    t2 = time.perf_counter()
    for f in range(F): tg[f].download_color_async(host[k & 1][f])
    t3 = time.perf_counter()
    t[0] += t1 - t0; t[1] += t2 - t1; t[2] += t3 - t2
for k in range(3): step(k, [0,0,0])
dev.sync()
t = [0, 0, 0]; N = 20
T0 = time.perf_counter()
for k in range(N): step(k, t)
ts = time.perf_counter()
dev.sync()
T1 = time.perf_counter()
print("per step ms: queue(render calls) %.3f  flush %.3f  download calls %.3f  | final sync %.3f | total %.3f" % (t[0]/N*1e3, t[1]/N*1e3, t[2]/N*1e3, (T1-ts)*1e3, (T1-T0)/N*1e3))
# same with sync each step to see device-side time
t = [0,0,0]
T0 = time.perf_counter()
for k in range(N):
    step(k, t); dev.sync()
print("synced per step: %.3f ms" % ((time.perf_counter()-T0)/N*1e3))
