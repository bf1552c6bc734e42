"""Regenerates the committed fixtures from the reference tree (run in the build container only).

    python tests/golden/make_golden.py [/root/reference]

Outputs (small, compressed; the reference tree does not exist on the GPU box):
  tests/golden/textured_quad.npz   <- core/tests/textured_quad.ppm   (core/tests/rendering.rs:47-50)
  tests/golden/triangle_fp.npz     <- core/triangle.ppm              (core/examples/hello_tri.rs:47-56, `fp` build)
  retrofire_b200/assets/bunny.obj.gz, crate.ppm.gz   <- demos/assets (scene inputs for C2 / C3; data, not code)
"""
import gzip
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from retrofire_b200.pnm import read_ppm  # noqa: E402

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
gold = os.path.join(ROOT, "tests", "golden")
np.savez_compressed(os.path.join(gold, "textured_quad.npz"), rgb=read_ppm(os.path.join(ref, "core/tests/textured_quad.ppm")))
np.savez_compressed(os.path.join(gold, "triangle_fp.npz"), rgb=read_ppm(os.path.join(ref, "core/triangle.ppm")))
assets = os.path.join(ROOT, "retrofire_b200", "assets")
for name in ("bunny.obj", "crate.ppm"):
    with open(os.path.join(ref, "demos/assets", name), "rb") as src, gzip.GzipFile(os.path.join(assets, name + ".gz"), "wb", mtime=0) as dst:
        shutil.copyfileobj(src, dst)
print("ok")
