"""Pins the CPU oracle against the reference's own golden vectors and known-answer tests
(SURVEY §8c). CPU only."""
import os

import numpy as np
import pytest

import retrofire_b200 as rf
from retrofire_b200 import scenes

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def run_scene(oracle, sc):
    tgt = oracle.HostTarget(sc.w, sc.h, sc.fmt, sc.has_depth)
    if sc.clear:
        tgt.clear(sc.ctx.color_clear, sc.ctx.depth_clear)
    stats = rf.Stats()
    for d in sc.draws:
        stats += oracle.render(d, tgt)
    return tgt, stats


def test_textured_quad_whole_frame(oracle):
    """core/tests/rendering.rs:18-60 — corner pixels and whole-frame equality with textured_quad.ppm."""
    tgt, st = run_scene(oracle, scenes.textured_quad())
    img = tgt.host_color()
    assert tuple(img[0, 0]) == (0, 0, 0xFF)
    assert tuple(img[255, 0]) == (0x7F, 0, 0)
    assert tuple(img[0, 255]) == (0x7F, 0, 0)
    gold = np.load(os.path.join(GOLD, "textured_quad.npz"))["rgb"]
    assert np.array_equal(img, gold)
    assert st.frags.i == st.frags.o == 65536 and st.prims.o == 2


def test_hello_tri_center_pixels(oracle):
    """core/examples/hello_tri.rs:47-53."""
    tgt, st = run_scene(oracle, scenes.hello_tri(fp=False))
    img = tgt.host_color()
    assert tuple(img[240, 320]) == (114, 102, 128, 255)
    assert st.frags.i == st.frags.o == 51200
    tgt, st = run_scene(oracle, scenes.hello_tri(fp=True))
    assert tuple(tgt.host_color()[240, 320]) == (151, 128, 187, 255)


def test_hello_tri_fp_matches_committed_triangle_ppm(oracle):
    """core/triangle.ppm (written by the example, fp build): coverage identical, colour within
    1 LSB (powf is libm-dependent; SURVEY §7 hard part 8)."""
    tgt, _ = run_scene(oracle, scenes.hello_tri(fp=True))
    img = tgt.host_color()
    gold = np.load(os.path.join(GOLD, "triangle_fp.npz"))["rgb"]
    cov = img[:, :, 3] != 0
    assert int(cov.sum()) == 51200
    assert np.array_equal(cov, gold.any(axis=2) | cov & (gold.sum(axis=2) == 0))  # covered set equal
    assert np.array_equal(gold.any(axis=2), cov)
    diff = np.abs(img[:, :, :3].astype(int) - gold.astype(int))
    assert diff.max() <= 1
    assert (diff > 0).mean() < 0.01


def test_shared_edge_no_gaps_no_overdraw(oracle):
    """render/raster.rs:326-368."""
    pts = [(8.0, 0.0, 0.0), (0.0, 6.0, 0.0), (14.0, 10.0, 0.0), (20.0, 3.0, 0.0)]
    V = [np.array(list(p) + [0.0], dtype=np.float32) for p in pts]
    buf = np.zeros((10, 20), dtype=int)
    for tri in ([V[0], V[1], V[2]], [V[0], V[2], V[3]]):
        spans, _ = oracle.tri_fill_spans(np.stack(tri), persp_mask=1)
        for y, x0, x1 in spans:
            buf[y, x0:x1] += 1
    expected = """
00000001110000000000
00000011111111000000
00000111111111111100
00011111111111111111
00111111111111111110
01111111111111111100
00111111111111111000
00000111111111110000
00000000011111100000
00000000000011000000"""
    got = "".join("\n" + "".join(str(v) for v in row) for row in buf)
    assert got == expected


def test_gradient_kat(oracle):
    """render/raster.rs:371-401."""
    verts = np.array([[15.0, 2.0, 1.0, 0.0], [2.0, 8.0, 1.0, 1.0], [26.0, 14.0, 1.0, 0.5]], dtype=np.float32)
    spans, vals = oracle.tri_fill_spans(verts, persp_mask=1)
    expected = """
              0
            2110
          3322211
       55444332221
     76665544433222
   88877666554443322
    98887766655444332
        88776665544433
            77666554443
                66655444
                    65544
                        54
"""
    s, k = "\n", 0
    for y, x0, x1 in spans:
        s += " " * max(x0, 1) if x0 else " "   # `{:w$}` of " " pads to width x0 (min 1 char)
        n = x1 - x0
        s += "".join(str(int(np.uint8(min(max(int(np.float32(10.0) * vals[k + i]), 0), 255)))) for i in range(n))
        s += "\n"
        k += n
    assert s == expected


def test_scanline_fragments_perspective_kat(oracle):
    """render/raster.rs:404-436: perspective-correct z and varying along one span."""
    w0, w1 = np.float32(2.0), np.float32(4.0)
    # Flat-bottom triangle whose scanline at y=42 runs 8..16 is awkward to build; check the recurrence directly:
    z0, z1 = np.float32(1) / w0, np.float32(1) / w1
    v0, v1 = np.float32(3.0) / w0, np.float32(5.0) / w1
    dz, dv = (z1 - z0) * (np.float32(1) / np.float32(8)), (v1 - v0) * (np.float32(1) / np.float32(8))
    zs = [2.0, 2.1333334, 2.2857144, 2.4615386, 2.6666667, 2.909091, 3.2, 3.5555556, 4.0]
    vs = [3.0, 3.1333334, 3.2857144, 3.4615386, 3.6666667, 3.909091, 4.2000003, 4.555556, 5.0]
    z, v = z0, v0
    for ze, ve in zip(zs, vs):
        assert abs(float(np.float32(1) / z) - ze) < 1e-5 * ze
        assert abs(float(v / z) - ve) < 1e-5 * ve
        z, v = np.float32(z + dz), np.float32(v + dv)


def test_clip_outcodes_and_single_cases(oracle):
    """render/clip.rs:425-470 outcode KATs."""
    assert oracle.outcode([0, 0, 0, 1]) == 0
    assert oracle.outcode([1, 0, 0, 1]) == 0
    assert oracle.outcode([0, -1, 0, 1]) == 0
    assert oracle.outcode([0, 1, 1, 1]) == 0
    assert oracle.outcode([0, 0, -1.5, 1]) == 0x01
    assert oracle.outcode([0, 0, 2, 1]) == 0x02
    assert oracle.outcode([-2, 0, 0, 1]) == 0x04
    assert oracle.outcode([3, 0, 0, 1]) == 0x08
    assert oracle.outcode([0, -1.001, 0, 1]) == 0x10
    assert oracle.outcode([0, 2, 0, 1]) == 0x20
    assert oracle.outcode([2, 2, 2, 1]) == 0x2A
    # fully inside passes unchanged; fully outside vanishes
    tri = np.array([[0, 0, 0, 1], [0.5, 0, 0, 1], [0, 0.5, 0, 1]], dtype=np.float32)
    assert np.array_equal(oracle.clip_tri(tri)[0], tri)
    assert oracle.clip_tri(tri + np.array([3, 0, 0, 0], dtype=np.float32)).shape[0] == 0


def test_edge_clip_kats(oracle):
    """render/clip.rs:455-486 (edge_clip_inside / outside / in_out / out_in): FAR_PLANE against single edges."""
    FAR = 1
    assert oracle.edge_plane(FAR, [2, 0, -1, 1], [-1, 1, 1, 1]) is None          # inside (touching the plane): unchanged
    assert oracle.edge_plane(FAR, [2, 0, 1.5, 1], [-1, 1, 2, 1]) is None         # both outside: nothing to intersect
    assert np.array_equal(oracle.edge_plane(FAR, [2, 0, 0, 1], [-1, 1, 2, 1]), np.array([0.5, 0.5, 1.0, 1.0], np.float32))
    assert np.array_equal(oracle.edge_plane(FAR, [2, 0, 4, 1], [-1, 1, 0, 1]), np.array([-0.25, 0.75, 1.0, 1.0], np.float32))


def test_clip_exhaustive_lattice_histogram(oracle):
    """render/clip.rs:667-719: 5^9 triangles, all outputs in bounds, output-count histogram."""
    hist, bad = oracle.clip_lattice_histogram()
    assert bad == 0
    assert hist == [559754, 536199, 537942, 254406, 58368, 6264, 192, 0]


def test_sampler_kats(oracle):
    """render/tex.rs:381-418 (SamplerRepeatPot, SamplerClamp)."""
    tex = np.array([[[0xFF, 0, 0], [0, 0xFF, 0]], [[0, 0, 0xFF], [0xFF, 0xFF, 0]]], dtype=np.uint8)
    rp = lambda u, v: oracle.sample(tex, 1, u, v)[:3]
    cl = lambda u, v: oracle.sample(tex, 0, u, v)[:3]
    assert rp(-0.1, 0.0) == (0, 0xFF, 0)
    assert rp(0.0, -0.1) == (0, 0, 0xFF)
    assert rp(1.0, 0.0) == (0xFF, 0, 0)
    assert rp(0.0, 1.0) == (0xFF, 0, 0)
    assert rp(4.8, 0.2) == (0, 0xFF, 0)
    assert rp(0.2, 4.8) == (0, 0, 0xFF)
    assert cl(-1.0, 0.0) == (0xFF, 0, 0)
    assert cl(0.0, -1.0) == (0xFF, 0, 0)
    assert cl(1.5, 0.0) == (0, 0xFF, 0)
    assert cl(0.0, 1.5) == (0, 0, 0xFF)
    assert cl(1.5, 1.5) == (0xFF, 0xFF, 0)
    on = lambda u, v: oracle.sample(tex, 2, u, v)[:3]   # SamplerOnce, tex.rs:410-418
    assert on(0.0, 0.0) == (0xFF, 0, 0)
    assert on(0.5, 0.0) == (0, 0xFF, 0)
    assert on(0.0, 0.5) == (0, 0, 0xFF)
    assert on(0.5, 0.5) == (0xFF, 0xFF, 0)
    assert on(1.0, 0.0) == (0xEE, 0xEE, 0xEE) and on(0.0, 1.5) == (0xEE, 0xEE, 0xEE)   # outside: the reference's slice index panics
    assert on(-0.4, 0.0) == (0xFF, 0, 0)                                              # `as u32` saturates a negative coordinate to 0


def test_xorshift64_kats():
    """core/src/math/rand.rs:134-137 and :314-318 (used to regenerate the sprites scene)."""
    g = scenes.Xorshift64(11223344556677889900)
    assert [g.next_bits() for _ in range(3)] == [7782624861773764242, 6203733934162558527, 13009646309496342147]
    g = scenes.Xorshift64()
    got = [float(g.uniform(-1.0, 1.0)) for _ in range(3)]
    assert got == [float(np.float32(x)) for x in (0.19692874, -0.7686298, 0.91969657)]


def test_matrix_kats():
    """core/src/math/mat.rs:1843-1863: viewport maps NDC corners exactly."""
    from retrofire_b200 import mathx as mx
    vp = mx.viewport((20, 10), (620, 470))
    p = vp @ np.array([-1, -1, 0, 1], dtype=np.float32)
    q = vp @ np.array([1, 1, 0, 1], dtype=np.float32)
    assert (p[0], p[1]) == (20, 10) and (q[0], q[1]) == (620, 470)


def test_pixfmt_kats(oracle):
    """core/src/util/pixfmt.rs:152-205."""
    import ctypes as C
    lib = oracle.load()
    col = (C.c_uint8 * 4)(0x11, 0x22, 0x33, 0x44)
    assert lib.rfo_pack_pixel(rf.FMT_XRGB8888, col) == 0x00112233
    assert lib.rfo_pack_pixel(rf.FMT_RGBA8888, col).to_bytes(4, "little") == bytes([0x11, 0x22, 0x33, 0x44])
    assert lib.rfo_pack_pixel(rf.FMT_ARGB8888, col).to_bytes(4, "little") == bytes([0x44, 0x11, 0x22, 0x33])
    assert lib.rfo_pack_pixel(rf.FMT_BGRA8888, col).to_bytes(4, "little") == bytes([0x33, 0x22, 0x11, 0x44])
    assert lib.rfo_pack_pixel(rf.FMT_RGBA4444, col) == 0x1234
    col = (C.c_uint8 * 4)(0x40, 0x20, 0x10, 0xFF)
    assert lib.rfo_pack_pixel(rf.FMT_RGB565, col) == 0b01000_001000_00010


def test_line_primitives_restatement_sanity(oracle):
    """raster::line (render/raster.rs:122-177) has no reference test: sanity of the restatement on axis-aligned and
    diagonal lines drawn through the full edge path (Render for Edge, prim.rs:41-60). Pinned by restatement only."""
    from retrofire_b200 import mathx as mx
    w, h = 32, 24
    ident = mx.identity()
    vp = mx.viewport((0, 0), (w, h))            # NDC (-1,-1) -> (0,0), (1,1) -> (w,h)

    def draw(p0, p1):
        verts = np.array([[p0[0], p0[1], 0, 1, 0, 0], [p1[0], p1[1], 0, 0, 1, 0]], dtype=np.float32)
        call = rf.DrawCall.make([[0, 1]], verts, rf.shader.new(rf.VS_MVP, rf.FS_COLOR3F), ident, vp, rf.Context(), edges=True)
        tgt = oracle.HostTarget(w, h, rf.FMT_RGBA8888, True)
        st = oracle.render(call, tgt)
        return (tgt.color != 0), st

    to_ndc = lambda x, y: (2 * x / w - 1, 2 * y / h - 1)
    cov, st = draw(to_ndc(4.0, 10.2), to_ndc(20.0, 10.2))     # horizontal: pixel centres 4.5 .. 19.5 on row 10
    assert st.prims.o == 1 and st.verts.o == 3 and st.frags.i == st.frags.o == 16
    assert cov[10, 4:20].all() and cov.sum() == 16
    cov, st = draw(to_ndc(7.3, 2.0), to_ndc(7.3, 12.0))       # vertical: rows 2 .. 11 in column 7
    assert cov[2:12, 7].all() and cov.sum() == 10
    cov, st = draw(to_ndc(2.0, 2.0), to_ndc(12.0, 12.0))      # diagonal: dx.abs() > dy is false -> tall branch, one pixel per row
    assert cov.sum() == 10 and all(cov[y].sum() == 1 for y in range(2, 12))
    cov, st = draw(to_ndc(5.0, 5.0), to_ndc(5.0, 5.0))        # zero length: nothing
    assert cov.sum() == 0 and st.prims.o == 1


@pytest.mark.parametrize("order", [rf.DepthSort.BackToFront, rf.DepthSort.FrontToBack])
def test_depth_sort_restatement_equals_presorted_submission(oracle, order):
    """Context::depth_sort (render.rs:180-182, 209-219) has no reference test. Without a depth test the painted result
    depends on the order of every overlap: rendering with depth_sort must equal rendering the same triangles submitted
    in an order computed independently here (numpy f32 restatement of Render::depth, prim.rs:21-23; stable on ties)."""
    import dataclasses
    from retrofire_b200 import scenes
    f32 = np.float32
    ctx = rf.Context(depth_test=None, face_cull=None)
    sc = scenes.random_soup(600, 256, 192, seed=5, lanes_kind="color3", big=True, clipy=False, ctx=ctx)   # no triangle is clipped
    d = sc.draws[0]
    m = np.asarray(d.uniform, f32).ravel()[:16].reshape(4, 4)     # VS_MVP: row-major Mat4 in the first 16 uniform floats
    v = d.verts[:, :3].astype(f32)
    clip = []
    for r in range(4):                                  # Mat4::apply: every row a left fold from 0.0, each step rounded to f32
        acc = f32(0) + m[r, 0] * v[:, 0]
        acc = acc + m[r, 1] * v[:, 1]
        acc = acc + m[r, 2] * v[:, 2]
        clip.append(acc + m[r, 3] * f32(1))
    cx, cy, cz, cw = clip
    inside = (np.abs(cx) < cw) & (np.abs(cy) < cw) & (np.abs(cz) < cw)
    t = d.prims.reshape(-1, 3)
    t = np.ascontiguousarray(t[inside[t].all(axis=1)])  # keep the triangles that are not clipped: their depth is that of the input
    assert len(t) > 200
    d = dataclasses.replace(d, prims=t)
    depth = ((cz[t[:, 0]] + cz[t[:, 1]]) + cz[t[:, 2]]) / f32(3)
    perm = np.argsort(depth if order == rf.DepthSort.FrontToBack else -depth, kind="stable")
    assert (np.diff(perm) != 1).any()                   # the sort really reorders this scene

    def paint(call):
        tgt = oracle.HostTarget(sc.w, sc.h, sc.fmt, True)
        tgt.clear(ctx.color_clear, ctx.depth_clear)
        st = oracle.render(call, tgt)
        return tgt.color.copy(), tgt.depth.copy(), st

    c_sorted, z_sorted, st_sorted = paint(dataclasses.replace(d, depth_sort=int(order)))
    c_manual, z_manual, st_manual = paint(dataclasses.replace(d, prims=np.ascontiguousarray(t[perm])))
    c_plain, _, _ = paint(d)
    assert (c_sorted == c_manual).all() and (z_sorted.view(np.uint32) == z_manual.view(np.uint32)).all()
    assert st_sorted.counters() == st_manual.counters()
    assert (c_sorted != c_plain).any()


def nan_max_scenes():
    """Two one-triangle scenes whose shaders see NaN in a `max` (also rendered on the device by tests/test_gpu_3_adversarial.py)."""
    from retrofire_b200 import mathx as mx
    f32 = np.float32
    pos = np.array([[-1, 1, 0], [1, 1, 0], [0, -1, 0]], f32)
    vp = mx.viewport((0, 32), (32, 0))
    ctx = rf.Context(face_cull=None, depth_test=None)
    tris = np.array([[0, 1, 2]], np.uint32)
    # crates shader: NaN normals, uv = 0 -> texel (0, 0) = (200, 100, 50) scaled by exactly 0.4
    tex = np.zeros((4, 4, 3), np.uint8); tex[0, 0] = (200, 100, 50)
    attr = np.concatenate([np.full((3, 3), np.nan, f32), np.zeros((3, 2), f32)], 1)
    shd = rf.shader.new(rf.VS_MVP, rf.FS_TEX_CLAMP_LIT, fs_uniform=[0.0, 0.0, -1.0], texture=rf.Texture(tex))
    lit = scenes.Scene("nanmax", 32, 32, rf.FMT_RGBA8888, False, ctx, [rf.DrawCall.make(tris, np.concatenate([pos, attr], 1), shd, np.eye(4, dtype=f32), vp, ctx)])
    # solids vertex shader: NaN spin matrix -> diffuse = 0.2 * 0.8, colour = ((n + 1.1) * 0.45) * diffuse with n = 0
    uni = (np.eye(4, dtype=f32), np.full((4, 4), np.nan, f32))
    shd = rf.shader.new(rf.VS_SOLIDS, rf.FS_COLOR3F)
    solids = scenes.Scene("nanmax2", 32, 32, rf.FMT_RGBA8888, False, ctx, [rf.DrawCall.make(tris, np.concatenate([pos, np.zeros((3, 3), f32)], 1), shd, uni, vp, ctx)])
    return lit, solids


def test_shader_max_ignores_nan_like_f32_max(oracle):
    """`f32::max` returns the other argument when one is NaN (Rust std), so `n.dot(&light_dir).max(0.0)` (crates.rs:44) with a NaN
    normal gives kd = lerp(0, 0.4, 1.0) = 0.4, and `(norm.z() + 0.2).max(0.2)` (solids.rs:75) gives 0.2 — not NaN (black), which is
    what C's std::max would produce. Found by tests/test_gpu_3_adversarial.py::test_lattice_shader_ties (zero-width first rows make
    dv_dx = 0 * inf = NaN)."""
    f32 = np.float32
    lit, solids = nan_max_scenes()
    tgt, st = run_scene(oracle, lit)
    img = tgt.host_color()
    want = [int(f32(256) * ((f32(c) / f32(256)) * f32(0.4))) for c in (200, 100, 50)] + [255]
    assert st.frags.o > 100 and img[16, 16].tolist() == want, (img[16, 16], want)
    tgt, st = run_scene(oracle, solids)
    c = (f32(0) + f32(1.1)) * f32(0.45) * (f32(0.2) * f32(0.8))
    assert tgt.host_color()[16, 16].tolist() == [int(f32(256) * c)] * 3 + [255], tgt.host_color()[16, 16]
