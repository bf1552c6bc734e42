"""Host-side mirror of retrofire-core's render API over the C ABI (include/retrofire_b200.h).

Mirrors, name for name where Python allows:

* `render(prims, verts, shader, uniform, to_screen, target, ctx)`  core/src/render.rs:134-147
* `Batch` (builder, `.render()`)                                   core/src/render/batch.rs:31-147
* `Context` (+ `FaceCull`, depth test, write masks, `stats`)       core/src/render/ctx.rs:11-127
* `Stats` / `Throughput`                                           core/src/render/stats.rs:16-40
* `shader.new(vs, fs)` -> a (vertex, fragment) catalogue pair      core/src/render/shader.rs:78-126
* `Texture`                                                        core/src/render/tex.rs:33-37
* `Framebuf` / `Colorbuf` targets (device-resident)                core/src/render/target.rs:35-58
* `Frame.clear`                                                    front/src/lib.rs:103-120

User closures cannot cross the FFI, so shaders are (vs_id, fs_id) pairs from the fixed
catalogue. Everything here is thin marshalling: the work happens in
librf_b200.so. There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _ffi
from ._ffi import RfDraw, RfStats


class RetrofireError(RuntimeError):
    """A non-zero rf_status (the reference panics in these cases)."""

    def __init__(self, status: int, msg: str = ""):
        self.status = status
        super().__init__(f"{_ffi.STATUS_NAMES.get(status, status)}: {msg}")


# ---- Stats (render/stats.rs) ------------------------------------------------------------
@dataclass
class Throughput:
    i: int = 0
    o: int = 0

    def __iadd__(self, other: "Throughput"):
        self.i += other.i
        self.o += other.o
        return self


@dataclass
class Stats:
    time: float = 0.0  # seconds
    calls: float = 0.0
    frames: float = 0.0
    objs: Throughput = field(default_factory=Throughput)
    prims: Throughput = field(default_factory=Throughput)
    verts: Throughput = field(default_factory=Throughput)
    frags: Throughput = field(default_factory=Throughput)

    def __iadd__(self, o: "Stats"):  # stats.rs:199-208
        self.time += o.time
        self.calls += o.calls
        self.frames += o.frames
        self.objs += o.objs
        self.prims += o.prims
        self.verts += o.verts
        self.frags += o.frags
        return self

    @staticmethod
    def from_c(s: RfStats) -> "Stats":
        return Stats(time=s.time_ns * 1e-9, calls=float(s.calls), objs=Throughput(s.objs_i, s.objs_o),
                     prims=Throughput(s.prims_i, s.prims_o), verts=Throughput(s.verts_i, s.verts_o),
                     frags=Throughput(s.frags_i, s.frags_o))

    def counters(self) -> tuple:
        return (int(self.calls), self.prims.i, self.prims.o, self.verts.i, self.verts.o, self.frags.i, self.frags.o)


# ---- Context (render/ctx.rs) ------------------------------------------------------------
class FaceCull:
    Front = _ffi.CULL_FRONT
    Back = _ffi.CULL_BACK


class DepthSort:  # render/ctx.rs:67-72
    FrontToBack = 1
    BackToFront = 2


class Ordering:  # core::cmp::Ordering as used by Context.depth_test
    Less = _ffi.DEPTH_LESS
    Equal = _ffi.DEPTH_EQUAL
    Greater = _ffi.DEPTH_GREATER


@dataclass
class Context:
    color_clear: Optional[tuple] = (0, 0, 0, 0xFF)
    depth_clear: Optional[float] = float("inf")
    face_cull: Optional[int] = FaceCull.Back
    depth_sort: Optional[int] = None
    depth_test: Optional[int] = Ordering.Less
    color_write: bool = True
    depth_write: bool = True
    stats: Stats = field(default_factory=Stats)


# ---- shaders (render/shader.rs; catalogue SURVEY §8a-11) ---------------------------------
_FS_LANES = {  # fs id -> (lanes, persp_mask) implied by the varying's Rust type
    _ffi.FS_COLOR3F: (3, 0b000), _ffi.FS_COLOR3F_SRGB: (3, 0b000), _ffi.FS_COLOR4F: (4, 0b0000),
    _ffi.FS_CHECKER: (2, 0b11), _ffi.FS_TEX_CLAMP_LIT: (5, 0b11111), _ffi.FS_TEX_CLAMP: (2, 0b11),
    _ffi.FS_TEX_REPEAT_POT: (2, 0b11), _ffi.FS_SPRITE_DISC: (2, 0b11), _ffi.FS_NORMAL_VIS: (3, 0b111),
    _ffi.FS_TEX_ONCE: (2, 0b11),
}


@dataclass
class Shader:
    vs: int
    fs: int
    lanes: int
    persp_mask: int
    fs_uniform: Sequence[float] = ()
    texture: Optional["Texture"] = None


class shader:  # namespace mirroring `render::shader`
    @staticmethod
    def new(vs: int, fs: int, *, fs_uniform: Sequence[float] = (), texture: "Texture" = None,
            lanes: int = None, persp_mask: int = None) -> Shader:
        dl, dm = _FS_LANES[fs]
        return Shader(vs, fs, dl if lanes is None else lanes, dm if persp_mask is None else persp_mask,
                      tuple(fs_uniform), texture)


# ---- Texture (render/tex.rs:33-37) --------------------------------------------------------
class Texture:
    """Host texel data (h, w, 3|4) uint8; uploaded lazily per Device."""

    def __init__(self, data: np.ndarray):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        assert data.ndim == 3 and data.shape[2] in (3, 4)
        self.data = data
        self.h, self.w = data.shape[:2]
        self.fmt = _ffi.TEXEL_RGB888 if data.shape[2] == 3 else _ffi.TEXEL_RGBA8888
        self._dev = {}

    def handle(self, dev: "Device") -> int:
        h = self._dev.get(dev.token)  # one token per Device instance, never reused: a closed Device's handles cannot be hit again
        if h is None:
            out = C.c_void_p()
            dev._check(dev.lib.rf_texture_create(dev.h, self.w, self.h, self.fmt, self.data.ctypes.data, self.w, C.byref(out)))
            h = out.value
            self._dev[dev.token] = h
            dev._textures.append(h)
            dev._texture_objs.append(self)
        return h


# ---- one render() call, marshalled ---------------------------------------------------------
def _flatten_uniform(uniform) -> np.ndarray:
    if isinstance(uniform, (tuple, list)):
        flat = np.concatenate([np.asarray(u, dtype=np.float32).reshape(-1) for u in uniform])
    else:
        flat = np.asarray(uniform, dtype=np.float32).reshape(-1)
    assert flat.size <= _ffi.RF_VS_UNIFORM_F32
    out = np.zeros(_ffi.RF_VS_UNIFORM_F32, dtype=np.float32)
    out[: flat.size] = flat
    return out


@dataclass(frozen=True, eq=False)
class DrawCall:
    """Arguments of one `render()` call in ABI form. Keeps the numpy buffers alive. Immutable (the marshalled rf_draw is
    memoised per Device): a changed uniform or array is a new DrawCall — `dataclasses.replace(d, uniform=...)`."""
    prims: np.ndarray      # (n,3) uint32
    verts: np.ndarray      # (n, stride) float32 : [x,y,z,a0..]
    shader: Shader
    uniform: np.ndarray    # 32 f32
    viewport: np.ndarray   # 4x4 f32
    face_cull: int = _ffi.CULL_BACK
    depth_test: int = _ffi.DEPTH_LESS
    color_write: bool = True
    depth_write: bool = True
    depth_sort: int = 0
    mesh: Optional["Mesh"] = None
    prim_kind: int = _ffi.PRIM_TRIS   # PRIM_EDGES: prims is (n,2) — `Edge<usize>` line primitives
    bbox: Optional[np.ndarray] = None  # (2,3) BBox<Model> low/upp: the draw is skipped when BBox::visibility(uniform) is Hidden

    def __post_init__(self):
        # `uniform` is copied into rf_draw.vs_uniform as RF_VS_UNIFORM_F32 floats: a shorter array (a caller replacing the
        # uniform of a two-matrix shader by one 4x4 matrix) is zero-padded here instead of being over-read there
        u = self.uniform
        if not (isinstance(u, np.ndarray) and u.dtype == np.float32 and u.size == _ffi.RF_VS_UNIFORM_F32 and u.flags.c_contiguous):
            object.__setattr__(self, "uniform", _flatten_uniform(u))
        object.__setattr__(self, "_cache", {})

    @staticmethod
    def make(prims, verts, shd: Shader, uniform, to_screen, ctx: Context = None, mesh: "Mesh" = None, edges: bool = False) -> "DrawCall":
        ctx = ctx or Context()
        if mesh is None:
            prims = np.ascontiguousarray(np.asarray(prims, dtype=np.uint32).reshape(-1, 2 if edges else 3))
            verts = np.ascontiguousarray(np.asarray(verts, dtype=np.float32))
            assert verts.ndim == 2 and verts.shape[1] >= 3 + shd.lanes, (verts.shape, shd.lanes)
        return DrawCall(prims, verts, shd, _flatten_uniform(uniform), np.asarray(to_screen, dtype=np.float32).reshape(4, 4),
                        face_cull=ctx.face_cull or 0, depth_test=ctx.depth_test or 0,
                        color_write=bool(ctx.color_write), depth_write=bool(ctx.depth_write),
                        depth_sort=0 if ctx.depth_sort is None else int(ctx.depth_sort), mesh=mesh,
                        prim_kind=_ffi.PRIM_EDGES if edges else _ffi.PRIM_TRIS)

    def cached_struct(self, key, texture_handle, mesh_handle) -> RfDraw:
        """to_struct() memoised per device: the marshalling costs more than the rf_render call."""
        c = self._cache
        st = c.get(key)
        if st is None:
            st = c[key] = self.to_struct(texture_handle, mesh_handle)
        return st

    def to_struct(self, texture_handle: int = None, mesh_handle: int = None) -> RfDraw:
        d = RfDraw()
        if self.mesh is None:
            d.indices = self.prims.ctypes.data_as(C.POINTER(C.c_uint32))
            d.n_prims = self.prims.shape[0]
            d.verts = self.verts.ctypes.data_as(C.POINTER(C.c_float))
            d.n_verts = self.verts.shape[0]
            d.vert_stride_f32 = self.verts.shape[1]
        else:
            d.mesh = mesh_handle
        d.n_attr_lanes = self.shader.lanes
        d.persp_mask = self.shader.persp_mask
        d.vs, d.fs = self.shader.vs, self.shader.fs
        C.memmove(d.vs_uniform, self.uniform.ctypes.data, 4 * _ffi.RF_VS_UNIFORM_F32)
        fu = np.zeros(_ffi.RF_FS_UNIFORM_F32, dtype=np.float32)
        fu[: len(self.shader.fs_uniform)] = np.asarray(self.shader.fs_uniform, dtype=np.float32)
        C.memmove(d.fs_uniform, fu.ctypes.data, 4 * _ffi.RF_FS_UNIFORM_F32)
        d.texture = texture_handle
        vp = np.ascontiguousarray(self.viewport, dtype=np.float32)
        C.memmove(d.viewport, vp.ctypes.data, 64)
        d.face_cull, d.depth_test = self.face_cull, self.depth_test
        d.color_write, d.depth_write, d.depth_sort = int(self.color_write), int(self.depth_write), self.depth_sort
        d.prim_kind = self.prim_kind
        if self.bbox is not None:
            d.bbox_cull = 1
            bb = np.ascontiguousarray(self.bbox, dtype=np.float32).reshape(6)
            C.memmove(d.bbox, bb.ctypes.data, 24)
        return d


def _peer_call(fn, head, world, rank, table):
    assert len(table) == world
    if isinstance(table[rank], (bytes, bytearray)):
        blob = (C.c_uint8 * (64 * world)).from_buffer_copy(b"".join(bytes(t) for t in table))
        return fn(*head, world, rank, blob, None)
    ptrs = (C.c_void_p * world)(*[int(t) for t in table])
    return fn(*head, world, rank, None, ptrs)


# ---- device objects -----------------------------------------------------------------------
class Device:
    """One rf_ctx (one GPU, one stream)."""

    def __init__(self, device: int = 0, stream: int = None):
        self.lib = _ffi.load()
        out = C.c_void_p()
        st = self.lib.rf_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(out))
        if st != _ffi.RF_OK:
            raise RetrofireError(st, "rf_ctx_create failed (is an sm_100 GPU visible?)")
        self.h = out.value
        self.token = object()   # identity of this Device for the caches of Texture / DrawCall (id() values are reused)
        self._textures = []
        self._texture_objs = []
        self._targets = []
        self._meshes = []
        self._pinned = []
        self._many = {}

    def _check(self, st: int):
        if st != _ffi.RF_OK:
            msg = self.lib.rf_last_error(self.h)
            raise RetrofireError(st, msg.decode() if msg else "")

    def close(self):
        if self.h:
            for t in self._targets:
                t._destroy()
            for m in self._meshes:
                m._destroy()
            for t in self._textures:
                self.lib.rf_texture_destroy(t)
            for tx in self._texture_objs:
                tx._dev.pop(self.token, None)
            self._many.clear()
            self.lib.rf_ctx_destroy(self.h)
            for p in self._pinned:
                self.lib.rf_host_free(p)
            self._pinned = []
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_geometry_async(self, on: bool = True):
        """Page-locked geometry arrays (pinned_empty) are read after render() returns: leave them alone until flush()/sync()."""
        self._check(self.lib.rf_ctx_set_geometry_async(self.h, int(on)))

    def set_row_band(self, y0: int, y1: int):
        self._check(self.lib.rf_ctx_set_row_band(self.h, y0, y1))

    # ---- sort-first over NVLink peer memory (rf_ctx_peer_* / rf_target_peer_*) -------------------
    def peer_export(self, ipc: bool = True):
        """This GPU's barrier slots: a 64-byte CUDA IPC handle (other processes) or the raw device pointer (same process)."""
        h, p = (C.c_uint8 * 64)(), C.c_void_p()
        self._check(self.lib.rf_ctx_peer_export(self.h, h if ipc else None, C.byref(p)))
        return bytes(h) if ipc else p.value

    def peer_attach(self, world: int, rank: int, table):
        """`table`: `world` IPC handles (bytes) or `world` device pointers (ints), in rank order."""
        self._check(_peer_call(self.lib.rf_ctx_peer_attach, (self.h,), world, rank, table))

    def replays(self) -> int:
        n = C.c_uint64()
        self._check(self.lib.rf_ctx_replays(self.h, C.byref(n)))
        return n.value

    def flush(self):
        self._check(self.lib.rf_flush(self.h))

    def sync(self):
        self._check(self.lib.rf_sync(self.h))

    def stats(self, reset: bool = False) -> Stats:
        s = RfStats()
        self._check(self.lib.rf_ctx_stats(self.h, C.byref(s), int(reset)))
        return Stats.from_c(s)

    def last_pass(self):
        t, n = C.c_uint64(), C.c_uint32()
        self._check(self.lib.rf_ctx_last_pass(self.h, C.byref(t), C.byref(n)))
        return t.value, n.value

    def profile(self, level: int = 2):
        """0 off; 1 time k_raster only (pass keeps its stream overlap); 2 time every kernel (serialised)."""
        self._check(self.lib.rf_ctx_profile(self.h, int(level)))

    def kernel_times(self) -> dict:
        """{kernel: (total_ns, launches)} since the last call (profiling mode)."""
        ns = (C.c_uint64 * _ffi.RF_N_KERNELS)()
        ln = (C.c_uint64 * _ffi.RF_N_KERNELS)()
        self._check(self.lib.rf_ctx_kernel_times(self.h, ns, ln))
        return {self.lib.rf_kernel_name(i).decode(): (int(ns[i]), int(ln[i])) for i in range(_ffi.RF_N_KERNELS)}

    def pinned_empty(self, shape, dtype) -> np.ndarray:
        """numpy array over rf_host_alloc'ed (page-locked) memory; freed with the Device."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        st = self.lib.rf_host_alloc(n, C.byref(p))
        if st != _ffi.RF_OK:
            raise RetrofireError(st, "rf_host_alloc")
        self._pinned.append(p.value)
        buf = (C.c_uint8 * n).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def framebuf(self, w: int, h: int, fmt: int = _ffi.FMT_RGBA8888, depth: bool = True) -> "Framebuf":
        return Framebuf(self, w, h, fmt, depth)

    def mesh(self, prims, verts, edges: bool = False) -> "Mesh":
        return Mesh(self, prims, verts, edges)

    # -- the hot path
    def render(self, call: DrawCall, target: "Framebuf", want_stats: bool = False) -> Optional[Stats]:
        tex = call.shader.texture.handle(self) if call.shader.texture is not None else None
        d = call.cached_struct(self.token, tex, call.mesh.h if call.mesh is not None else None)
        if want_stats:
            s = RfStats()
            self._check(self.lib.rf_render(self.h, target.h, C.byref(d), C.byref(s)))
            return Stats.from_c(s)
        self._check(self.lib.rf_render(self.h, target.h, C.byref(d), None))
        return None

    def render_many(self, calls: Sequence[DrawCall], target: "Framebuf"):
        """All render() calls of one frame into one target with a single crossing of the C ABI (rf_render_many). The marshalled
        array is memoised on the list object: a frame loop that re-submits the same list pays for the marshalling once."""
        memo = self._many.get(id(calls))
        if memo is None or memo[0] is not calls or memo[1] != len(calls) or any(a is not b for a, b in zip(memo[3], calls)):
            arr = (RfDraw * len(calls))()
            keep = []
            for i, call in enumerate(calls):
                tex = call.shader.texture.handle(self) if call.shader.texture is not None else None
                d = call.to_struct(tex, call.mesh.h if call.mesh is not None else None)
                C.memmove(C.byref(arr, i * C.sizeof(RfDraw)), C.byref(d), C.sizeof(RfDraw))
                keep.append(call)
            memo = (calls, len(calls), arr, keep)
            # the memo keeps the list (and through it the caller's arrays) alive: bound it, oldest entries first, so that a loop
            # which builds a new list every frame does not grow it without end (a frame batch re-submits a few hundred lists)
            while len(self._many) >= 4096:
                self._many.pop(next(iter(self._many)))
            self._many[id(calls)] = memo
        self._check(self.lib.rf_render_many(self.h, target.h, memo[2], memo[1]))

    def render_frames(self, call: DrawCall, targets: Sequence["Framebuf"], uniforms: np.ndarray):
        """Frame batch: draw i -> targets[i] with vs_uniform uniforms[i] (n, 32) f32."""
        uniforms = np.ascontiguousarray(uniforms, dtype=np.float32).reshape(len(targets), _ffi.RF_VS_UNIFORM_F32)
        tex = call.shader.texture.handle(self) if call.shader.texture is not None else None
        d = call.to_struct(tex, call.mesh.h if call.mesh is not None else None)
        arr = (C.c_void_p * len(targets))(*[t.h for t in targets])
        self._check(self.lib.rf_render_frames(self.h, arr, len(targets), C.byref(d), uniforms.ctypes.data))


class Mesh:
    """Persistent device copy of (prims, verts) — avoids Batch's per-call clones (batch.rs:62-84)."""

    def __init__(self, dev: Device, prims, verts, edges: bool = False):
        self.dev = dev
        prims = np.ascontiguousarray(np.asarray(prims, dtype=np.uint32).reshape(-1, 2 if edges else 3))
        verts = np.ascontiguousarray(np.asarray(verts, dtype=np.float32))
        self.n_prims, self.n_verts, self.stride = prims.shape[0], verts.shape[0], verts.shape[1]
        out = C.c_void_p()
        dev._check(dev.lib.rf_mesh_create(dev.h, verts.ctypes.data, self.n_verts, self.stride, prims.ctypes.data,
                                          self.n_prims, _ffi.PRIM_EDGES if edges else _ffi.PRIM_TRIS, C.byref(out)))
        self.h = out.value
        dev._meshes.append(self)

    def _destroy(self):
        if self.h:
            self.dev.lib.rf_mesh_destroy(self.h)
            self.h = None


_HOST_DTYPE = {_ffi.FMT_RGBA8888: (np.uint8, 4), _ffi.FMT_XRGB8888: (np.uint32, 1), _ffi.FMT_ARGB8888: (np.uint8, 4),
               _ffi.FMT_BGRA8888: (np.uint8, 4), _ffi.FMT_RGB888: (np.uint8, 3), _ffi.FMT_RGB565: (np.uint8, 2),
               _ffi.FMT_RGBA4444: (np.uint8, 2)}


class Framebuf:
    """Device-resident `Framebuf<Colorbuf<_, Fmt>, Buf2<f32>>` (or a bare colour buffer if depth=False)."""

    def __init__(self, dev: Device, w: int, h: int, fmt: int, depth: bool):
        self.dev, self.w, self.h_px, self.fmt, self.has_depth = dev, w, h, fmt, depth
        out = C.c_void_p()
        dev._check(dev.lib.rf_target_create(dev.h, w, h, fmt, int(depth), C.byref(out)))
        self.h = out.value
        dev._targets.append(self)

    def _destroy(self):
        if self.h:
            self.dev.lib.rf_target_destroy(self.h)
            self.h = None

    def clear(self, ctx: Context = None):
        """Frame::clear (front/src/lib.rs:103-120): depth is filled with 1/depth_clear."""
        ctx = ctx or Context()
        rgba = (C.c_uint8 * 4)(*ctx.color_clear) if ctx.color_clear is not None else None
        z = None
        if ctx.depth_clear is not None and self.has_depth:
            z = C.c_float(float(np.float32(1.0) / np.float32(ctx.depth_clear)))
        self.dev._check(self.dev.lib.rf_target_clear(self.dev.h, self.h, rgba, C.byref(z) if z is not None else None))

    def peer_export(self, ipc: bool = True):
        h, p = (C.c_uint8 * 64)(), C.c_void_p()
        self.dev._check(self.dev.lib.rf_target_peer_export(self.dev.h, self.h, h if ipc else None, C.byref(p)))
        return bytes(h) if ipc else p.value

    def peer_attach(self, world: int, rank: int, table):
        """Replicate this GPU's colour stores into the same target on the other GPUs (handles or pointers in rank order)."""
        self.dev._check(_peer_call(self.dev.lib.rf_target_peer_attach, (self.dev.h, self.h), world, rank, table))

    def _host_shape(self):
        dt, ch = _HOST_DTYPE[self.fmt]
        return (dt, (self.h_px, self.w) if ch == 1 else (self.h_px, self.w, ch))

    def download_color(self) -> np.ndarray:
        dt, shape = self._host_shape()
        out = np.empty(shape, dtype=dt)
        self.dev._check(self.dev.lib.rf_target_download_color(self.dev.h, self.h, out.ctypes.data, self.w))
        return out

    def upload_color(self, buf: np.ndarray):
        dt, shape = self._host_shape()
        buf = np.ascontiguousarray(buf, dtype=dt).reshape(shape)
        self.dev._check(self.dev.lib.rf_target_upload_color(self.dev.h, self.h, buf.ctypes.data, self.w))

    def download_color_async(self, out: np.ndarray):
        """Queue a D2H copy into `out` (ideally from Device.pinned_empty); valid after Device.sync()."""
        self.dev._check(self.dev.lib.rf_target_download_color_async(self.dev.h, self.h, out.ctypes.data, self.w))

    def download_depth(self) -> np.ndarray:
        out = np.empty((self.h_px, self.w), dtype=np.float32)
        self.dev._check(self.dev.lib.rf_target_download_depth(self.dev.h, self.h, out.ctypes.data, self.w))
        return out

    def upload_depth(self, buf: np.ndarray):
        buf = np.ascontiguousarray(buf, dtype=np.float32).reshape(self.h_px, self.w)
        self.dev._check(self.dev.lib.rf_target_upload_depth(self.dev.h, self.h, buf.ctypes.data, self.w))

    def color_devptr(self) -> int:
        return self.dev.lib.rf_target_color_devptr(self.h)

    def depth_devptr(self) -> int:
        return self.dev.lib.rf_target_depth_devptr(self.h)


# ---- render() and Batch ----------------------------------------------------------------------
def render(prims, verts, shd: Shader, uniform, to_screen, target: Framebuf, ctx: Context, *, sync_stats: bool = True):
    """`retrofire_core::render::render` (render.rs:134-207).

    With sync_stats=True (default, the reference's observable behaviour) the call returns
    after the draw has executed and `ctx.stats` has been updated (render.rs:206). With
    sync_stats=False the draw is only queued; stats accumulate in the Device (`Device.stats`).
    """
    call = DrawCall.make(prims, verts, shd, uniform, to_screen, ctx)
    s = target.dev.render(call, target, want_stats=sync_stats)
    if s is not None:
        ctx.stats += s


@dataclass
class Batch:
    """`render::Batch` (batch.rs:31-147): builder whose `.render()` calls `render()`."""
    prims: np.ndarray = None
    verts: np.ndarray = None
    uniform_: object = None
    shader_: Shader = None
    viewport_: np.ndarray = None
    target_: Framebuf = None
    ctx: Context = field(default_factory=Context)

    def _upd(self, **kw) -> "Batch":
        return dataclasses.replace(self, **kw)

    def primitives(self, prims): return self._upd(prims=np.array(prims, dtype=np.uint32).reshape(-1, 3))
    def vertices(self, verts): return self._upd(verts=np.array(verts, dtype=np.float32))
    def mesh(self, m): return self._upd(prims=np.array(m[0], dtype=np.uint32).reshape(-1, 3), verts=np.array(m[1], dtype=np.float32))
    def uniform(self, u): return self._upd(uniform_=u)
    def shader(self, s): return self._upd(shader_=s)
    def viewport(self, v): return self._upd(viewport_=v)
    def target(self, t): return self._upd(target_=t)
    def context(self, c): return self._upd(ctx=c)
    def clone(self): return self._upd()

    def render(self, *, sync_stats: bool = True):
        render(self.prims, self.verts, self.shader_, self.uniform_, self.viewport_, self.target_, self.ctx, sync_stats=sync_stats)
