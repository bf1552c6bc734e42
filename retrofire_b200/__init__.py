"""retrofire_b200 — B200-native (sm_100a) implementation of retrofire's render() hot path.

The product is `librf_b200.so` (hand-written CUDA behind the C ABI in
include/retrofire_b200.h); this package is the thin host-side mirror of retrofire-core's
render API on top of it. No CPU fallback: using a Device without the built library or
without a GPU raises.
"""
from . import _ffi, mathx, scene, scenes, text  # noqa: F401
from ._ffi import *  # noqa: F401,F403  (enum constants)
from .api import (Batch, Context, DepthSort, Device, DrawCall, FaceCull, Framebuf, Mesh, Ordering, RetrofireError, Shader, Stats,  # noqa: F401
                  Texture, Throughput, render, shader)

__version__ = "0.1.0"
