"""Small scenes through every kernel path, for compute-sanitizer (memcheck / racecheck / initcheck):
   compute-sanitizer --tool memcheck python scratch/sanitize.py"""
import sys, dataclasses
sys.path.insert(0, '/root/repo')
import numpy as np
import retrofire_b200 as rf
from retrofire_b200 import scenes
from tests.parity import run_gpu, run_oracle, assert_parity
from oracle import rfo
dev = rf.Device(0)
nod = dict(depth_test=None, face_cull=None)
S = [scenes.hello_tri(), scenes.textured_quad(),
     scenes.random_soup(600, 320, 200, seed=1, lanes_kind="color3", big=False),
     scenes.random_soup(200, 320, 200, seed=2, lanes_kind="lit", big=True),
     scenes.random_soup(300, 320, 200, seed=3, lanes_kind="lanes8", big=True),
     scenes.random_soup(1500, 96, 64, seed=4, lanes_kind="color3", big=True),          # every tile heaviest
     scenes.random_lines(400, 320, 200, seed=5),
     scenes.random_soup(500, 320, 200, seed=6, lanes_kind="color3", big=True, ctx=rf.Context(depth_sort=rf.DepthSort.BackToFront, **nod)),
     scenes.bunny(subdiv=0, w=400, h=300), scenes.hello_text(0.7, w=400, h=300),
     scenes.crates("169", 480, 270, device_cull=True), scenes.sprites(500, w=320, h=200)]
quick = len(sys.argv) > 1
for sc in (S[:4] if quick else S):
    want = run_oracle(rfo, sc)
    for rep in range(2):
        assert_parity(run_gpu(dev, sc), want, name=f"{sc.name}-{rep}", color_tol=0)
    print("ok", sc.name, flush=True)
dev.close()
print("done")
