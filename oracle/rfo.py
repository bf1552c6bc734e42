"""ctypes binding of the CPU oracle (oracle/rf_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs may import this module. The product (retrofire_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from retrofire_b200 import _ffi
from retrofire_b200._ffi import RfDraw, RfStats
from retrofire_b200.api import DrawCall, Stats

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "librf_oracle.so")


class RfoTarget(C.Structure):
    _fields_ = [("w", C.c_uint32), ("h", C.c_uint32), ("fmt", C.c_uint32), ("color", C.c_void_p),
                ("depth", C.c_void_p), ("band_y0", C.c_uint32), ("band_y1", C.c_uint32)]


class RfoTexture(C.Structure):
    _fields_ = [("w", C.c_uint32), ("h", C.c_uint32), ("fmt", C.c_uint32), ("data", C.c_void_p), ("stride", C.c_uint64)]


class SpanRec(C.Structure):
    _fields_ = [("y", C.c_uint64), ("x0", C.c_uint64), ("x1", C.c_uint64)]


_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "rf_oracle.cpp")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-s"] + (["-B"] if force else []), check=True)
    return LIB_PATH


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = C.CDLL(LIB_PATH)
        lib.rfo_render.restype = C.c_int
        lib.rfo_render.argtypes = [C.POINTER(RfDraw), C.POINTER(RfoTexture), C.POINTER(RfoTarget), C.POINTER(RfStats)]
        lib.rfo_clip_tri.restype = C.c_int
        lib.rfo_clip_tri.argtypes = [C.c_void_p, C.c_void_p]
        lib.rfo_clip_lattice_histogram.restype = C.c_int64
        lib.rfo_clip_lattice_histogram.argtypes = [C.c_void_p]
        lib.rfo_outcode.restype = C.c_uint8
        lib.rfo_outcode.argtypes = [C.c_void_p]
        lib.rfo_edge_plane.restype = C.c_int
        lib.rfo_edge_plane.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.rfo_tri_fill_spans.restype = C.c_int
        lib.rfo_tri_fill_spans.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.POINTER(SpanRec), C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        lib.rfo_sample.restype = None
        lib.rfo_sample.argtypes = [C.POINTER(RfoTexture), C.c_int, C.c_float, C.c_float, C.c_void_p]
        lib.rfo_pack_pixel.restype = C.c_uint32
        lib.rfo_pack_pixel.argtypes = [C.c_uint32, C.c_void_p]
        _lib = lib
    return _lib


class HostTarget:
    """Host Buf2 pair in the same uint32-container convention as the device targets."""

    def __init__(self, w: int, h: int, fmt: int = _ffi.FMT_RGBA8888, depth: bool = True):
        self.w, self.h, self.fmt = w, h, fmt
        self.color = np.zeros((h, w), dtype=np.uint32)  # Buf2::new zero-fills, util/buf.rs:155-161
        self.depth = np.zeros((h, w), dtype=np.float32) if depth else None
        self.band = (0, h)

    def clear(self, color_rgba=(0, 0, 0, 0xFF), depth_clear=float("inf")):
        """Frame::clear, front/src/lib.rs:103-120."""
        if color_rgba is not None:
            c = (C.c_uint8 * 4)(*color_rgba)
            self.color[:] = load().rfo_pack_pixel(self.fmt, c)
        if depth_clear is not None and self.depth is not None:
            self.depth[:] = np.float32(1.0) / np.float32(depth_clear)

    def host_color(self) -> np.ndarray:
        return container_to_host(self.fmt, self.color)


def container_to_host(fmt: int, cont: np.ndarray) -> np.ndarray:
    """uint32 containers -> the host element layout rf_target_download_color produces."""
    h, w = cont.shape
    by = np.ascontiguousarray(cont).view(np.uint8).reshape(h, w, 4)
    if fmt in (_ffi.FMT_RGBA8888, _ffi.FMT_ARGB8888, _ffi.FMT_BGRA8888):
        return by.copy()
    if fmt == _ffi.FMT_XRGB8888:
        return cont.copy()
    if fmt == _ffi.FMT_RGB888:
        return by[:, :, :3].copy()
    return by[:, :, :2].copy()  # 565 / 4444: native-endian u16


def render(call: DrawCall, target: HostTarget) -> Stats:
    """Run one render() through the oracle; returns this call's Stats. Raises on non-zero status."""
    lib = load()
    d = call.to_struct(None, None)
    assert call.mesh is None, "oracle takes host geometry"
    tex = None
    if call.shader.texture is not None:
        t = call.shader.texture
        tex = RfoTexture(t.w, t.h, t.fmt, t.data.ctypes.data, t.w)
    tg = RfoTarget(target.w, target.h, target.fmt, target.color.ctypes.data,
                   target.depth.ctypes.data if target.depth is not None else None, target.band[0], target.band[1])
    s = RfStats()
    st = lib.rfo_render(C.byref(d), C.byref(tex) if tex is not None else None, C.byref(tg), C.byref(s))
    if st != 0:
        from retrofire_b200.api import RetrofireError
        raise RetrofireError(st, "oracle")
    return Stats.from_c(s)


def clip_tri(pos: np.ndarray) -> np.ndarray:
    pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(12)
    out = np.zeros(84, dtype=np.float32)
    n = load().rfo_clip_tri(pos.ctypes.data, out.ctypes.data)
    return out[: 12 * n].reshape(n, 3, 4)


def clip_lattice_histogram():
    hist = np.zeros(8, dtype=np.int64)
    bad = load().rfo_clip_lattice_histogram(hist.ctypes.data)
    return hist.tolist(), int(bad)


def outcode(pos) -> int:
    p = np.ascontiguousarray(pos, dtype=np.float32).reshape(4)
    return int(load().rfo_outcode(p.ctypes.data))


def edge_plane(plane: int, a, b):
    """ClipPlane::intersect of edge a-b with plane index `plane` (0 near, 1 far, 2 left, 3 right, 4 bottom, 5 top): position or None."""
    pa, pb = (np.ascontiguousarray(p, dtype=np.float32).reshape(4) for p in (a, b))
    out = np.zeros(4, dtype=np.float32)
    return out if load().rfo_edge_plane(plane, pa.ctypes.data, pb.ctypes.data, out.ctypes.data) else None


def tri_fill_spans(lanes: np.ndarray, persp_mask: int = 0, max_spans: int = 4096, max_frags: int = 1 << 20):
    lanes = np.ascontiguousarray(lanes, dtype=np.float32)
    L = lanes.shape[1] - 3
    spans = (SpanRec * max_spans)()
    fv = np.zeros(max_frags, dtype=np.float32)
    nf = C.c_int()
    n = load().rfo_tri_fill_spans(lanes.ctypes.data, L, persp_mask, spans, max_spans, fv.ctypes.data, max_frags, C.byref(nf))
    return [(s.y, s.x0, s.x1) for s in spans[:n]], fv[: nf.value]


def sample(tex_data: np.ndarray, kind: int, u: float, v: float):
    tex_data = np.ascontiguousarray(tex_data, dtype=np.uint8)
    h, w, ch = tex_data.shape
    t = RfoTexture(w, h, _ffi.TEXEL_RGB888 if ch == 3 else _ffi.TEXEL_RGBA8888, tex_data.ctypes.data, w)
    out = (C.c_uint8 * 4)()
    load().rfo_sample(C.byref(t), kind, u, v, out)
    return tuple(out)
