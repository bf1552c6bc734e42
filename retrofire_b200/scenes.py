"""Scene builders for the BASELINE.json configurations, as lists of DrawCalls.

Host-side asset/scene preparation — the step before the hot path (reference: demos/src/bin/*,
geom/src/solids/platonic.rs, geom/src/io.rs, core/src/math/rand.rs). These follow the
reference's scene definitions; the oracle and the GPU path consume the *same* arrays, so
parity does not depend on these being bit-identical to the Rust builders.
"""
from __future__ import annotations

import gzip
import math
import os
from dataclasses import dataclass, field
from typing import List

import numpy as np

from . import _ffi
from . import mathx as mx
from .api import Context, DrawCall, FaceCull, Shader, Texture, shader

f32 = np.float32
ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")


@dataclass
class Scene:
    name: str
    w: int
    h: int
    fmt: int
    has_depth: bool
    ctx: Context                      # clear colour / depth for Frame::clear
    draws: List[DrawCall] = field(default_factory=list)
    clear: bool = True                # False: bare zero-initialised Buf2 (no Frame::clear)

    @property
    def n_prims(self) -> int:
        return sum(d.prims.shape[0] for d in self.draws)

    @property
    def n_verts(self) -> int:
        return sum(d.verts.shape[0] for d in self.draws)

    def geometry_bytes(self) -> int:
        """B_geom of SURVEY §8d: n_verts*4*(3+A) + n_prims*12 per draw."""
        return sum(d.verts.shape[0] * 4 * (3 + d.shader.lanes) + d.prims.shape[0] * 12 for d in self.draws)


# ---- C1: core/examples/hello_tri.rs:5-45 ------------------------------------------------------
def hello_tri(fp: bool = False) -> Scene:
    verts = np.array([[-1, 1, 0, 1.0, 0.0, 0.0], [1, 1, 0, 0.0, 0.8, 0.0], [0, -1, 0, 0.4, 0.4, 1.0]], dtype=f32)
    w, h = 640, 480
    mvp = mx.then(mx.translate3(0, 0, 2), mx.perspective(1.0, f32(w) / f32(h), 0.1, 1000.0))
    vp = mx.viewport((0, h), (w, 0))
    shd = shader.new(_ffi.VS_MVP_LINEARIZE if fp else _ffi.VS_MVP, _ffi.FS_COLOR3F_SRGB if fp else _ffi.FS_COLOR3F)
    ctx = Context()
    call = DrawCall.make([[0, 1, 2]], verts, shd, mvp, vp, ctx)
    return Scene("hello_tri_fp" if fp else "hello_tri", w, h, _ffi.FMT_RGBA8888, False, ctx, [call], clear=False)


# ---- core/tests/rendering.rs:18-60 ---------------------------------------------------------------
def textured_quad() -> Scene:
    verts = np.array([[-1, -1, 0, 0, 0], [1, -1, 0, 0, 1], [-1, 1, 0, 1, 0], [1, 1, 0, 1, 1]], dtype=f32)
    faces = [[0, 1, 2], [3, 2, 1]]
    tex = np.zeros((8, 8, 4), dtype=np.uint8)
    for y in range(8):
        for x in range(8):
            xor = (x ^ y) & 1
            tex[y, x] = (0x7F * xor, 0, 0xFF * (1 - xor), 0)
    w = h = 256
    mvp = mx.then(mx.translate3(0, 0, 1), mx.perspective(1.0, 1.0, 0.1, 1000.0))
    vp = mx.viewport((0, 0), (w, h))
    shd = shader.new(_ffi.VS_MVP, _ffi.FS_TEX_CLAMP, texture=Texture(tex))
    ctx = Context()
    return Scene("textured_quad", w, h, _ffi.FMT_RGB888, False, ctx, [DrawCall.make(faces, verts, shd, mvp, vp, ctx)], clear=False)


# ---- mesh helpers ---------------------------------------------------------------------------------
def read_obj(path: str):
    """Minimal `v`/`f` reader (geom/src/io.rs:145-229): triangles and quads, 1-based indices."""
    opener = gzip.open if path.endswith(".gz") else open
    vs, fs = [], []
    with opener(path, "rt") as fh:
        for line in fh:
            t = line.split()
            if not t:
                continue
            if t[0] == "v":
                vs.append([float(t[1]), float(t[2]), float(t[3])])
            elif t[0] == "f":
                idx = [int(s.split("/")[0]) - 1 for s in t[1:]]
                fs.append(idx[:3])
                if len(idx) == 4:
                    fs.append([idx[0], idx[2], idx[3]])
    return np.array(vs, dtype=f32), np.array(fs, dtype=np.uint32)


def vertex_normals(pos: np.ndarray, faces: np.ndarray) -> np.ndarray:
    """Area-weighted vertex normals (core/src/geom/mesh.rs:205-233)."""
    a, b, c = pos[faces[:, 0]], pos[faces[:, 1]], pos[faces[:, 2]]
    fn = np.cross(b - a, c - a).astype(f32)
    n = np.zeros_like(pos)
    for k in range(3):
        np.add.at(n, faces[:, k], fn)
    ln = np.sqrt((n * n).sum(1, keepdims=True))
    ln[ln == 0] = 1
    return (n / ln).astype(f32)


def subdivide(pos: np.ndarray, faces: np.ndarray):
    """Midpoint (1->4) subdivision with shared edge midpoints."""
    e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]]).astype(np.int64)
    es = np.sort(e, axis=1)
    key = es[:, 0] * (pos.shape[0] + 1) + es[:, 1]
    uniq, inv = np.unique(key, return_inverse=True)
    first = np.zeros(uniq.shape[0], dtype=np.int64)
    first[inv] = np.arange(key.shape[0])
    mid = ((pos[es[first, 0]] + pos[es[first, 1]]) * f32(0.5)).astype(f32)
    npos = np.concatenate([pos, mid])
    F = faces.shape[0]
    m01, m12, m20 = (inv[:F] + pos.shape[0]), (inv[F:2 * F] + pos.shape[0]), (inv[2 * F:] + pos.shape[0])
    v0, v1, v2 = faces[:, 0].astype(np.int64), faces[:, 1].astype(np.int64), faces[:, 2].astype(np.int64)
    nf = np.stack([np.stack([v0, m01, m20], 1), np.stack([m01, v1, m12], 1), np.stack([m20, m12, v2], 1),
                   np.stack([m01, m12, m20], 1)], 1).reshape(-1, 3)
    return npos, nf.astype(np.uint32)


_bunny_cache = {}


def bunny_mesh(subdiv: int = 2):
    """demos/src/bin/solids.rs:192-199: scale 0.12, translate -Y, vertex normals; then `subdiv` x 1->4."""
    if subdiv not in _bunny_cache:
        pos, faces = read_obj(os.path.join(ASSETS, "bunny.obj.gz"))
        pos = (pos * f32(0.12) + np.array([0, -1, 0], dtype=f32)).astype(f32)
        for _ in range(subdiv):
            pos, faces = subdivide(pos, faces)
        nrm = vertex_normals(pos, faces)
        _bunny_cache[subdiv] = (np.concatenate([pos, nrm], 1).astype(f32), faces)
    return _bunny_cache[subdiv]


def solids_uniform(theta: float, w: int, h: int):
    """(mvp, spin) of demos/src/bin/solids.rs:60-64,98-108 at time `theta` (carousel idle)."""
    th = f32(theta)
    spin = mx.then(mx.rotate_x(f32(th * f32(0.37))), mx.rotate_y(f32(th * f32(0.51))))
    aspect = f32(w) / f32(h)
    proj = mx.perspective(mx.fov_equiv35mm(28.0), aspect, 0.1, 1000.0)
    w2p = mx.then(mx.scale3(1, -1, -1), proj)
    mvp = mx.then(mx.then(mx.then(spin, mx.translate3(0, 0, -3)), mx.identity()), w2p)
    return mvp, spin


# ---- C2: bunny, Gouraud + depth ---------------------------------------------------------------------
def bunny(subdiv: int = 2, theta: float = 1.0, w: int = 1920, h: int = 1080) -> Scene:
    verts, faces = bunny_mesh(subdiv)
    ctx = Context(color_clear=(0x33, 0x33, 0x33, 0xFF))
    vp = mx.viewport((10, h - 10), (w - 10, 10))
    shd = shader.new(_ffi.VS_SOLIDS, _ffi.FS_COLOR3F)
    call = DrawCall.make(faces, verts, shd, solids_uniform(theta, w, h), vp, ctx)
    return Scene(f"bunny_x{4 ** subdiv}", w, h, _ffi.FMT_XRGB8888, True, ctx, [call])


# ---- C3: crates ------------------------------------------------------------------------------------------
_BOX_COORDS = np.array([[0, 0, 0], [0, 0, 1], [0, 1, 0], [0, 1, 1], [1, 0, 0], [1, 0, 1], [1, 1, 0], [1, 1, 1]], dtype=f32)
_BOX_NORMS = np.array([[-1, 0, 0], [1, 0, 0], [0, -1, 0], [0, 1, 0], [0, 0, -1], [0, 0, 1]], dtype=f32)
_BOX_UV = np.array([[0, 0], [1, 0], [0, 1], [1, 1]], dtype=f32)
_BOX_VERTS = [(0b011, 0, 0), (0b010, 0, 1), (0b001, 0, 2), (0b000, 0, 3), (0b110, 1, 0), (0b111, 1, 1), (0b100, 1, 2), (0b101, 1, 3),
              (0b000, 2, 0), (0b100, 2, 1), (0b001, 2, 2), (0b101, 2, 3), (0b011, 3, 0), (0b111, 3, 1), (0b010, 3, 2), (0b110, 3, 3),
              (0b010, 4, 0), (0b110, 4, 1), (0b000, 4, 2), (0b100, 4, 3), (0b111, 5, 0), (0b011, 5, 1), (0b101, 5, 2), (0b001, 5, 3)]
_BOX_FACES = np.array([[0, 1, 3], [0, 3, 2], [4, 5, 7], [4, 7, 6], [8, 9, 11], [8, 11, 10], [12, 13, 15], [12, 15, 14],
                       [16, 17, 19], [16, 19, 18], [20, 21, 23], [20, 23, 22]], dtype=np.uint32)


def cube_mesh(side: float = 2.0):
    """geom/src/solids/platonic.rs:131-241: Cube{side_len} with (Normal3, TexCoord) attributes."""
    d = f32(side) / f32(2)
    lo, hi = np.array([-d] * 3, dtype=f32), np.array([d] * 3, dtype=f32)
    out = []
    for pi, ni, ti in _BOX_VERTS:
        p = lo + (hi - lo) * _BOX_COORDS[pi]
        out.append(np.concatenate([p, _BOX_NORMS[ni], _BOX_UV[ti]]))
    return np.array(out, dtype=f32), _BOX_FACES.copy()


def floor_mesh(size: int = 50):
    """demos/src/bin/crates.rs:156-191."""
    verts, faces = [], []
    wd = size * 2 + 1
    for j in range(-size, size + 1):
        for i in range(-size, size + 1):
            io, jo = i & 1, j & 1
            verts.append([i, -1.0, j, io, jo])
            if j > -size and i > -size:
                jj, ii = size + j, size + i
                a, b, c, d = wd * (jj - 1) + (ii - 1), wd * (jj - 1) + ii, wd * jj + (ii - 1), wd * jj + ii
                if io ^ jo:
                    faces += [[a, c, d], [a, d, b]]
                else:
                    faces += [[b, c, d], [b, a, c]]
    return np.array(verts, dtype=f32), np.array(faces, dtype=np.uint32)


def _outcodes(clip: np.ndarray) -> np.ndarray:
    x, y, z, w = clip[..., 0], clip[..., 1], clip[..., 2], clip[..., 3]
    oc = (-z - w > 0) * 1 + (z - w > 0) * 2 + (-x - w > 0) * 4 + (x - w > 0) * 8 + (-y - w > 0) * 16 + (y - w > 0) * 32
    return oc.astype(np.uint8)


def _bbox_hidden(verts: np.ndarray, m2p: np.ndarray) -> bool:
    """BBox::visibility == Hidden (render/scene.rs:81-87): all 8 corners outside one plane."""
    lo, hi = verts[:, :3].min(0), verts[:, :3].max(0)
    corners = np.array([[x, y, z, 1] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])], dtype=f32)
    oc = _outcodes((corners @ m2p.T).astype(f32))
    return bool(np.bitwise_and.reduce(oc) != 0)


def crate_texture() -> Texture:
    from .pnm import read_ppm
    return Texture(read_ppm(os.path.join(ASSETS, "crate.ppm.gz")))


def crates(grid: str = "1089", w: int = 3840, h: int = 2160, host_cull: bool = True, device_cull: bool = False) -> Scene:
    """demos/src/bin/crates.rs. grid "169": reference layout (-30..=30 step 5); "1089": -48..=48 step 3.
    host_cull: the demo's `bbox.visibility(..) == Hidden` test runs here and hidden objects are not submitted;
    device_cull: EVERY object is submitted with its BBox<Model> and the test runs on the device (rf_draw.bbox_cull)."""
    import dataclasses
    if device_cull:
        host_cull = False
    from .scene import BBox   # Obj::with_transform computes the box from the mesh (scene.rs:29-39)
    with_bbox = lambda call, verts: dataclasses.replace(call, bbox=BBox.of(verts).as_array()) if device_cull else call
    ctx = Context()
    vw, vh = w - 20, h - 20
    vp = mx.viewport((10, h - 10), (w - 10, 10))
    aspect = f32(vw) / f32(vh)
    proj = mx.perspective(mx.fov_diagonal(math.radians(90.0), aspect), aspect, 0.1, 1000.0)
    w2v = mx.mat([[0, 0, -1, 0], [0, 1, 0, 0], [1, 0, 0, 0], [0, 0, 0, 1]])  # FirstPerson::default, cam.rs:404-418
    w2p = mx.then(w2v, proj)
    light = mx.normalize([-2.0, 1.0, -4.0])
    tex = crate_texture()
    floor_shd = shader.new(_ffi.VS_MVP, _ffi.FS_CHECKER)
    crate_shd = shader.new(_ffi.VS_MVP, _ffi.FS_TEX_CLAMP_LIT, fs_uniform=light, texture=tex)
    draws = []
    fv, ff = floor_mesh(50)
    m2p = mx.then(mx.identity(), w2p)
    if not (host_cull and _bbox_hidden(fv, m2p)):
        draws.append(with_bbox(DrawCall.make(ff, fv, floor_shd, m2p, vp, ctx), fv))
    cv, cf = cube_mesh(2.0)
    rng = range(-30, 31, 5) if grid == "169" else range(-48, 49, 3)
    for i in rng:
        for j in rng:
            m2p = mx.then(mx.translate3(i, 0, j), w2p)
            if host_cull and _bbox_hidden(cv, m2p):
                continue
            draws.append(with_bbox(DrawCall.make(cf, cv, crate_shd, m2p, vp, ctx), cv))
    return Scene(f"crates_{grid}", w, h, _ffi.FMT_RGBA8888, True, ctx, draws)


# ---- C4: sprites ---------------------------------------------------------------------------------------------
class Xorshift64:
    """core/src/math/rand.rs:118-183."""
    DEFAULT_SEED = 378682147834061

    def __init__(self, seed: int = DEFAULT_SEED):
        assert seed != 0
        self.x = seed & 0xFFFFFFFFFFFFFFFF

    def next_bits(self) -> int:
        x = self.x
        x ^= (x << 13) & 0xFFFFFFFFFFFFFFFF
        x ^= x >> 7
        x ^= (x << 17) & 0xFFFFFFFFFFFFFFFF
        self.x = x
        return x

    def uniform(self, lo: float, hi: float) -> np.float32:
        """Uniform<f32>::sample, rand.rs:320-326."""
        bits = (127 << 23) | (self.next_bits() >> 41)
        unit = f32(np.array([bits], dtype=np.uint32).view(f32)[0] - f32(1.0))
        return f32(f32(unit * f32(f32(hi) - f32(lo))) + f32(lo))


def points_in_unit_ball(rng: Xorshift64, n: int) -> np.ndarray:
    """PointsInUnitBall (rand.rs:530-576): rejection sampling of Uniform([-1;3]..[1;3])."""
    out = np.empty((n, 3), dtype=f32)
    k = 0
    while k < n:
        v = np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-1, 1)], dtype=f32)
        if f32(f32(f32(v[0] * v[0]) + f32(v[1] * v[1])) + f32(v[2] * v[2])) <= 1.0:
            out[k] = v
            k += 1
    return out


def sprites(count: int = 10000, theta: float = 1.0, w: int = 1920, h: int = 1080) -> Scene:
    """demos/src/bin/sprites.rs:14-77."""
    quad = np.array([[-1, -1], [-1, 1], [1, -1], [1, 1]], dtype=f32)
    pts = points_in_unit_ball(Xorshift64(), count)
    verts = np.concatenate([np.repeat(pts, 4, axis=0), np.tile(quad, (count, 1))], axis=1).astype(f32)
    base = (4 * np.arange(count, dtype=np.uint32))[:, None]
    tris = np.stack([base + np.array([0, 1, 3], dtype=np.uint32), base + np.array([0, 3, 2], dtype=np.uint32)], 1).reshape(-1, 3)
    ctx = Context()
    aspect = f32(w) / f32(h)
    proj = mx.perspective(1.0, aspect, 1e-2, 1e3)
    th = f32(theta)
    mv = mx.then(mx.then(mx.rotate_x(f32(th * f32(0.2))), mx.rotate_z(f32(th * f32(0.14)))), mx.translate3(0, 0, 0.5))
    vp = mx.viewport((10, h - 10), (w - 10, 10))
    shd = shader.new(_ffi.VS_SPRITE, _ffi.FS_SPRITE_DISC)
    call = DrawCall.make(tris, verts, shd, (mv, proj), vp, ctx)
    return Scene(f"sprites_{count}", w, h, _ffi.FMT_XRGB8888, True, ctx, [call])


# ---- C5-i: N small independent triangles ---------------------------------------------------------------------------
def small_tris(n: int = 1_000_000, w: int = 7680, h: int = 4320, seed: int = 1, cull: bool = True) -> Scene:
    """circumradius 1–8 px, z~U[1,100], random winding; numpy generator seeded `seed`."""
    g = np.random.default_rng(seed)
    z = g.uniform(1, 100, n).astype(f32)
    cx = (g.uniform(-1, 1, n).astype(f32) * z).astype(f32)
    cy = (g.uniform(-1, 1, n).astype(f32) * z * f32(h / w)).astype(f32)
    r = (g.uniform(2, 16, n).astype(f32) / f32(w / 2) * z * f32(0.5)).astype(f32)
    ang = g.uniform(0, 2 * math.pi, (n, 3)).astype(f32)
    px = cx[:, None] + r[:, None] * np.cos(ang)
    py = cy[:, None] + r[:, None] * np.sin(ang)
    pz = np.repeat(z[:, None], 3, 1)
    col = g.uniform(0, 1, (n, 3, 3)).astype(f32)
    verts = np.concatenate([np.stack([px, py, pz], 2), col], 2).reshape(3 * n, 6).astype(f32)
    tris = np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
    ctx = Context(face_cull=FaceCull.Back if cull else None)
    mvp = mx.perspective(1.0, f32(w) / f32(h), 0.5, 1000.0)
    vp = mx.viewport((0, h), (w, 0))
    shd = shader.new(_ffi.VS_MVP, _ffi.FS_COLOR3F)
    return Scene(f"small_tris_{n}", w, h, _ffi.FMT_RGBA8888, True, ctx, [DrawCall.make(tris, verts, shd, mvp, vp, ctx)])


# ---- generic randomized scene for parity fuzzing ---------------------------------------------------------------------
def random_soup(n: int, w: int, h: int, seed: int, *, lanes_kind: str = "color3", big: bool = False, clipy: bool = True,
                ctx: Context = None) -> Scene:
    """Random triangles, many crossing the frustum planes (exercises Sutherland–Hodgman), mixed sizes."""
    g = np.random.default_rng(seed)
    ctx = ctx or Context()
    spread = 1.6 if clipy else 0.9
    z = g.uniform(-0.5 if clipy else 1.0, 30, (n, 1)).astype(f32)
    c = g.uniform(-spread, spread, (n, 1, 2)).astype(f32) * np.maximum(np.abs(z), 0.5)[:, :, None]
    rad = g.uniform(0.01, 1.5 if big else 0.15, (n, 1, 1)).astype(f32) * np.maximum(np.abs(z), 0.5)[:, :, None]
    xy = c + rad * g.uniform(-1, 1, (n, 3, 2)).astype(f32)
    zz = z[:, :, None] + g.uniform(-1, 1, (n, 3, 1)).astype(f32) * (2.0 if clipy else 0.2)
    pos = np.concatenate([xy, zz], 2).astype(f32)
    if lanes_kind == "color3":
        attr = g.uniform(0, 1, (n, 3, 3)).astype(f32)
        shd = shader.new(_ffi.VS_MVP, _ffi.FS_COLOR3F)
    elif lanes_kind == "uv":
        attr = g.uniform(-0.5, 1.5, (n, 3, 2)).astype(f32)
        tex = g.integers(0, 256, (16, 16, 3), dtype=np.uint8)
        shd = shader.new(_ffi.VS_MVP, _ffi.FS_TEX_REPEAT_POT, texture=Texture(tex))
    elif lanes_kind == "disc":
        attr = g.uniform(-1.2, 1.2, (n, 3, 2)).astype(f32)
        shd = shader.new(_ffi.VS_MVP, _ffi.FS_SPRITE_DISC)
    elif lanes_kind == "lit":
        nrm = g.normal(size=(n, 3, 3)).astype(f32)
        nrm /= np.linalg.norm(nrm, axis=2, keepdims=True)
        attr = np.concatenate([nrm, g.uniform(-0.2, 1.2, (n, 3, 2)).astype(f32)], 2).astype(f32)
        tex = g.integers(0, 256, (32, 32, 3), dtype=np.uint8)
        shd = shader.new(_ffi.VS_MVP, _ffi.FS_TEX_CLAMP_LIT, fs_uniform=mx.normalize([-2.0, 1.0, -4.0]), texture=Texture(tex))
    elif lanes_kind == "color4":
        attr = g.uniform(-0.1, 1.1, (n, 3, 4)).astype(f32)
        shd = shader.new(_ffi.VS_MVP, _ffi.FS_COLOR4F)
    elif lanes_kind == "checker":
        attr = g.uniform(-0.5, 1.5, (n, 3, 2)).astype(f32)
        shd = shader.new(_ffi.VS_MVP, _ffi.FS_CHECKER)
    elif lanes_kind == "normal":
        attr = g.normal(size=(n, 3, 3)).astype(f32)
        shd = shader.new(_ffi.VS_MVP, _ffi.FS_NORMAL_VIS)
    elif lanes_kind == "texclamp":
        attr = g.uniform(-0.5, 1.5, (n, 3, 2)).astype(f32)
        tex = g.integers(0, 256, (13, 7, 4), dtype=np.uint8)   # non-power-of-two RGBA texture
        shd = shader.new(_ffi.VS_MVP, _ffi.FS_TEX_CLAMP, texture=Texture(tex))
    elif lanes_kind == "lanes8":
        # 8 varying lanes (the widest kernel instantiation); the colour shader reads the first three,
        # two of which are declared perspective-divided to exercise z_div on a colour shader
        attr = g.uniform(0.05, 1, (n, 3, 8)).astype(f32)
        shd = shader.new(_ffi.VS_MVP, _ffi.FS_COLOR3F, lanes=8, persp_mask=0b10100101)
    else:
        raise ValueError(lanes_kind)
    verts = np.concatenate([pos, attr], 2).reshape(3 * n, -1).astype(f32)
    tris = np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
    mvp = mx.perspective(1.0, f32(w) / f32(h), 0.5, 50.0)
    vp = mx.viewport((0, h), (w, 0))
    return Scene(f"soup_{lanes_kind}_{n}_{seed}", w, h, _ffi.FMT_RGBA8888, True, ctx, [DrawCall.make(tris, verts, shd, mvp, vp, ctx)])


# ---- Edge primitives (render/prim.rs:41-60, raster.rs:122-177): wireframes ---------------------------------------------
def mesh_edges(faces: np.ndarray) -> np.ndarray:
    """Unique undirected edges of a triangle mesh, as (n,2) uint32 (what render/debug.rs draws as a wireframe)."""
    e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]]).astype(np.int64)
    e = np.unique(np.sort(e, axis=1), axis=0)
    return e.astype(np.uint32)


def bunny_wireframe(subdiv: int = 0, theta: float = 1.0, w: int = 1920, h: int = 1080) -> Scene:
    verts, faces = bunny_mesh(subdiv)
    ctx = Context(color_clear=(0x33, 0x33, 0x33, 0xFF))
    vp = mx.viewport((10, h - 10), (w - 10, 10))
    shd = shader.new(_ffi.VS_SOLIDS, _ffi.FS_COLOR3F)
    call = DrawCall.make(mesh_edges(faces), verts, shd, solids_uniform(theta, w, h), vp, ctx, edges=True)
    return Scene(f"bunny_wire_x{4 ** subdiv}", w, h, _ffi.FMT_XRGB8888, True, ctx, [call])


def random_lines(n: int, w: int, h: int, seed: int, ctx: Context = None, lanes_kind: str = "color3") -> Scene:
    """Random line segments, many crossing the frustum planes; all slopes, including axis-aligned and degenerate ones."""
    g = np.random.default_rng(seed)
    ctx = ctx or Context()
    z = g.uniform(-0.5, 30, (n, 1)).astype(f32)
    c = g.uniform(-1.5, 1.5, (n, 1, 2)).astype(f32) * np.maximum(np.abs(z), 0.5)[:, :, None]
    rad = g.uniform(0.0, 1.2, (n, 1, 1)).astype(f32) * np.maximum(np.abs(z), 0.5)[:, :, None]
    xy = c + rad * g.uniform(-1, 1, (n, 2, 2)).astype(f32)
    k = n // 10
    xy[:k, 1, 1] = xy[:k, 0, 1]            # horizontal
    xy[k:2 * k, 1, 0] = xy[k:2 * k, 0, 0]  # vertical
    xy[2 * k:2 * k + 5, 1] = xy[2 * k:2 * k + 5, 0]  # zero length
    zz = z[:, :, None] + g.uniform(-1, 1, (n, 2, 1)).astype(f32) * 2.0
    pos = np.concatenate([xy, zz], 2).astype(f32)
    if lanes_kind == "color3":
        attr = g.uniform(0, 1, (n, 2, 3)).astype(f32)
        shd = shader.new(_ffi.VS_MVP, _ffi.FS_COLOR3F)
    else:
        attr = g.uniform(-0.5, 1.5, (n, 2, 2)).astype(f32)
        tex = g.integers(0, 256, (16, 16, 3), dtype=np.uint8)
        shd = shader.new(_ffi.VS_MVP, _ffi.FS_TEX_REPEAT_POT, texture=Texture(tex))
    verts = np.concatenate([pos, attr], 2).reshape(2 * n, -1).astype(f32)
    edges = np.arange(2 * n, dtype=np.uint32).reshape(n, 2)
    mvp = mx.perspective(1.0, f32(w) / f32(h), 0.5, 50.0)
    vp = mx.viewport((0, h), (w, 0))
    return Scene(f"lines_{lanes_kind}_{n}_{seed}", w, h, _ffi.FMT_RGBA8888, True, ctx, [DrawCall.make(edges, verts, shd, mvp, vp, ctx, edges=True)])


# ---- demos/src/bin/hello.rs: text as textured geometry (render/text.rs) ---------------------------------------------
def synthetic_font(sub_dims=(16, 24), glyphs=256, seed=7) -> "text.Atlas":
    """A 16-column grid atlas of black/white glyph cells (the demo's font is a PBM bitmap; this one is procedural:
    every cell is a distinct blocky pattern with a one-texel dark border so neighbouring glyphs are distinguishable)."""
    from . import text
    gw, gh = sub_dims
    cols = 16
    rows = (glyphs + cols - 1) // cols
    g = np.random.default_rng(seed)
    tex = np.zeros((rows * gh, cols * gw, 3), np.uint8)
    for i in range(glyphs):
        cell = np.kron(g.integers(0, 2, ((gh - 2) // 2, (gw - 2) // 2)), np.ones((2, 2), np.int64))
        y0, x0 = i // cols * gh, i % cols * gw
        tex[y0 + 1:y0 + 1 + cell.shape[0], x0 + 1:x0 + 1 + cell.shape[1]] = (cell * 255)[:, :, None]
    return text.Atlas(sub_dims, Texture(tex))


def hello_text(secs: float = 0.7, msg: str = "   Hello,\nRetrocomputing\n     World!", w: int = 800, h: int = 600) -> Scene:
    """hello.rs:13-73: the message as glyph quads, SamplerClamp into the font atlas, no face culling, viewport inset by 10."""
    from . import text
    t = text.Text(synthetic_font()).write(msg)
    faces, verts = t.geom
    vp_m = mx.then(mx.translate3(0, 0, 15.0), mx.perspective(1.0, f32(4.0) / f32(3.0), 0.1, 1000.0))
    mvp = mx.then(mx.then(mx.then(mx.then(mx.scale3(0.1, 0.1, 0.1), mx.translate3(-10.0, -5.0, f32(5.0) * f32(math.sin(secs)))),
                                  mx.rotate_y(f32(secs) * f32(0.59))), mx.rotate_z(f32(math.sin(secs * 1.13)))), vp_m)
    ctx = Context(face_cull=None)
    shd = shader.new(_ffi.VS_MVP, _ffi.FS_TEX_CLAMP, texture=t.font.texture)
    call = DrawCall.make(faces, verts, shd, mvp, mx.viewport((10, 10), (w - 10, h - 10)), ctx)
    return Scene("hello_text", w, h, _ffi.FMT_RGBA8888, True, ctx, [call])
