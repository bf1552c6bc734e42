"""GPU parity, part 1 (collected first): the BASELINE.json configurations at full size (hello_tri, bunny native and x16, crates
169 / 1,089 cubes at 4K, sprites 10 k, 1 M small triangles at 8K), the frame batch the bench times, the Context flag matrix, every
pixel format, the shader catalogue, lines, object culling, depth sort and row bands — the CUDA path through the C ABI against the
CPU oracle on identical inputs. Bit-exact depth (coverage and depth-test winners), bit-exact colour except where a shader uses powf
(±1 LSB), equal Stats counters. The adversarial families live in test_gpu_3_adversarial.py so that one failure there
cannot hide these."""
import dataclasses
import os

import numpy as np
import pytest

import retrofire_b200 as rf
from retrofire_b200 import _ffi, scenes
from tests.parity import assert_parity, depth_equal, run_gpu, run_oracle

f32 = np.float32

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def check(device, oracle, sc, **kw):
    assert_parity(run_gpu(device, sc), run_oracle(oracle, sc), name=sc.name, **kw)


def test_hello_tri(device, oracle):
    """BASELINE config 1 (core/examples/hello_tri.rs), non-fp shaders: bit-exact."""
    sc = scenes.hello_tri(fp=False)
    got = run_gpu(device, sc)
    assert tuple(got[0][240, 320]) == (114, 102, 128, 255)
    assert_parity(got, run_oracle(oracle, sc), name=sc.name)


def test_hello_tri_fp_and_golden(device, oracle):
    """fp shaders use powf: coverage exact, colour within 1 LSB of oracle and of core/triangle.ppm."""
    sc = scenes.hello_tri(fp=True)
    got = run_gpu(device, sc)
    assert_parity(got, run_oracle(oracle, sc), name=sc.name, color_tol=1)
    gold = np.load(os.path.join(GOLD, "triangle_fp.npz"))["rgb"]
    assert np.array_equal(got[0][:, :, 3] != 0, gold.any(axis=2))
    assert np.abs(got[0][:, :, :3].astype(int) - gold.astype(int)).max() <= 1
    assert np.abs(got[0][240, 320].astype(int) - np.array([151, 128, 187, 255])).max() <= 1


def test_textured_quad_golden(device, oracle):
    """core/tests/rendering.rs: whole frame equals textured_quad.ppm."""
    sc = scenes.textured_quad()
    got = run_gpu(device, sc)
    gold = np.load(os.path.join(GOLD, "textured_quad.npz"))["rgb"]
    assert np.array_equal(got[0], gold)
    assert_parity(got, run_oracle(oracle, sc), name=sc.name)


def test_bunny_native(device, oracle):
    """BASELINE config 2 at the reference asset's size (4,968 tris), 1920x1080."""
    check(device, oracle, scenes.bunny(subdiv=0))


def test_bunny_x16(device, oracle):
    """BASELINE config 2: ~79k triangles."""
    check(device, oracle, scenes.bunny(subdiv=2))


def test_crates_169_reduced(device, oracle):
    """BASELINE config 3, reference layout, 1920x1080: 170 draws, long perspective-correct spans."""
    check(device, oracle, scenes.crates("169", w=1920, h=1080))


def test_crates_1089_full_4k(device, oracle):
    """BASELINE config 3 at full size: 3840x2160, one draw per cube + floor, perspective-correct textured."""
    check(device, oracle, scenes.crates("1089"))


def test_sprites(device, oracle):
    """BASELINE config 4 (reduced count for CPU time): discard + heavy overdraw."""
    check(device, oracle, scenes.sprites(count=3000))


def test_sprites_10k_full(device, oracle):
    """BASELINE config 4 at full size: 10,000 sphere sprites (20,000 tris), discard + overdraw, 1920x1080."""
    check(device, oracle, scenes.sprites(10000))


def test_small_tris_1m_8k(device, oracle):
    """BASELINE config 5(i) at full size: 1,000,000 small triangles at 7680x4320."""
    check(device, oracle, scenes.small_tris(1_000_000))


def test_frame_batch_render_frames(device, oracle):
    """rf_render_frames: one mesh, per-frame uniforms, per-frame targets (SURVEY 8e frame sharding)."""
    verts, faces = scenes.bunny_mesh(0)
    frames = [scenes.bunny(subdiv=0, theta=0.5 * f, w=640, h=360) for f in range(4)]
    mesh = device.mesh(faces, verts)
    import dataclasses
    call = dataclasses.replace(frames[0].draws[0], mesh=mesh)
    targets = [device.framebuf(640, 360, frames[0].fmt, True) for _ in frames]
    for t in targets:
        t.clear(frames[0].ctx)
    device.stats(reset=True)
    device.render_frames(call, targets, np.stack([f.draws[0].uniform for f in frames]))
    got_stats = device.stats(reset=True)
    want_total = rf.Stats()
    for f, t in zip(frames, targets):
        wc, wd, ws = run_oracle(oracle, f)
        want_total += ws
        assert np.array_equal(t.download_color(), wc) and depth_equal(t.download_depth(), wd)
    assert got_stats.counters() == want_total.counters()


@pytest.mark.parametrize("ctxkw", [
    dict(face_cull=None), dict(face_cull=rf.FaceCull.Front), dict(depth_test=None),
    dict(depth_test=rf.Ordering.Greater), dict(depth_test=rf.Ordering.Equal),
    dict(color_write=False), dict(depth_write=False), dict(depth_test=None, depth_write=False),
])
def test_context_flags(device, oracle, ctxkw):
    """Context fields consumed by the path (render/ctx.rs:11-101); no reference test pins these."""
    ctx = rf.Context(**ctxkw)
    if ctxkw.get("depth_test") in (rf.Ordering.Greater, rf.Ordering.Equal):
        ctx.depth_clear = 0.001  # 1/depth_clear = 1000: 'Greater' passes when curr > new
    sc = scenes.random_soup(1500, 400, 300, seed=11, lanes_kind="color3", big=True, ctx=ctx)
    check(device, oracle, sc)


@pytest.mark.parametrize("fmt", [rf.FMT_RGBA8888, rf.FMT_XRGB8888, rf.FMT_ARGB8888, rf.FMT_BGRA8888, rf.FMT_RGB888,
                                 rf.FMT_RGB565, rf.FMT_RGBA4444])
def test_pixel_formats(device, oracle, fmt):
    """util/pixfmt.rs conversions at write time and in download."""
    sc = scenes.random_soup(300, 256, 128, seed=3, lanes_kind="color3", big=True)
    sc.fmt = fmt
    check(device, oracle, sc)


def test_multi_draw_frame_and_per_draw_stats(device, oracle):
    """Several render() calls into one target (front/src/minifb.rs frame loop), mixed shaders."""
    a = scenes.random_soup(800, 512, 384, seed=21, lanes_kind="lit", big=True)
    b = scenes.random_soup(800, 512, 384, seed=22, lanes_kind="color3", big=False)
    c = scenes.random_soup(800, 512, 384, seed=23, lanes_kind="disc", big=True)
    a.draws += b.draws + c.draws
    a.name = "multi"
    want = run_oracle(oracle, a)
    assert_parity(run_gpu(device, a), want, name="multi-queued")
    assert_parity(run_gpu(device, a, per_draw_sync=True), want, name="multi-sync")


@pytest.mark.parametrize("kind", ["color3", "uv", "disc", "lit"])
@pytest.mark.parametrize("big", [False, True])
def test_random_soup(device, oracle, kind, big):
    """Random triangles crossing all six frustum planes; small and screen-filling; every lane layout."""
    for seed in (1, 2):
        sc = scenes.random_soup(3000 if not big else 400, 640, 360, seed=seed, lanes_kind=kind, big=big)
        check(device, oracle, sc)


@pytest.mark.parametrize("kind", ["color4", "checker", "normal", "texclamp", "lanes8"])
def test_remaining_catalogue_shaders_and_widest_lanes(device, oracle, kind):
    """FS_COLOR4F, FS_CHECKER, FS_NORMAL_VIS, FS_TEX_CLAMP on a non-POT RGBA texture, and 8 varying lanes."""
    for big in (False, True):
        check(device, oracle, scenes.random_soup(1200 if not big else 300, 512, 300, seed=13, lanes_kind=kind, big=big))


@pytest.mark.parametrize("order", [rf.DepthSort.BackToFront, rf.DepthSort.FrontToBack])
@pytest.mark.parametrize("dtest", [None, rf.Ordering.Less])
def test_depth_sort(device, oracle, order, dtest):
    """Context::depth_sort (render.rs:180-182, 209-219): clipped primitives sorted by Render::depth before
    rasterisation. Without a depth test (painter's algorithm) every overlap depends on the order; triangles crossing
    the frustum planes are sorted by the depth of their clipped fan pieces."""
    ctx = rf.Context(depth_sort=order, depth_test=dtest, face_cull=None)
    for big in (False, True):
        sc = scenes.random_soup(3000 if not big else 400, 640, 360, seed=31, lanes_kind="color3", big=big, ctx=ctx)
        check(device, oracle, sc)
        if dtest is None:   # the order matters in this scene: the unsorted submission paints something else
            import dataclasses
            plain = dataclasses.replace(sc, draws=[dataclasses.replace(sc.draws[0], depth_sort=0)])
            assert (run_gpu(device, plain)[0] != run_gpu(device, sc)[0]).any()


@pytest.mark.parametrize("kind", ["color3", "uv"])
def test_line_primitives(device, oracle, kind):
    """Edge<usize> primitives: Render for Edge, Clip for [Edge], raster::line — all slopes, clipped
    against every frustum plane, axis-aligned and zero-length segments."""
    for seed in (1, 2):
        check(device, oracle, scenes.random_lines(3000, 640, 360, seed=seed, lanes_kind=kind))


def test_wireframe_over_solid_and_front_cull(device, oracle):
    """A wireframe pass over the solid mesh in one frame (lines + triangles, depth tested), and FaceCull::Front,
    which culls every edge because Render::is_backface defaults to false (render.rs:72-74, ctx.rs:95-101)."""
    solid = scenes.bunny(subdiv=0, w=960, h=540)
    wire = scenes.bunny_wireframe(subdiv=0, w=960, h=540)
    solid.draws = solid.draws + wire.draws
    check(device, oracle, solid)
    culled = scenes.random_lines(500, 320, 240, seed=3, ctx=rf.Context(face_cull=rf.FaceCull.Front))
    got = run_gpu(device, culled)
    assert got[2].prims.o == 0 and got[2].frags.i == 0
    assert_parity(got, run_oracle(oracle, culled), name="front-cull-lines")


def test_object_culling_on_the_device(device, oracle):
    """the scene loop of crates.rs:100-131 with `BBox::visibility` (scene.rs:81-87) evaluated on the device.
    All 170 objects are submitted with their bounding boxes; hidden ones are skipped as if render() had not been called
    (Stats count only the rendered ones, objs.i/objs.o as the demo counts them), and the frame equals the host-culled one."""
    dev_sc = scenes.crates("169", 960, 540, device_cull=True)
    host_sc = scenes.crates("169", 960, 540)
    assert len(dev_sc.draws) == 170 and len(host_sc.draws) < 100
    got, want = run_gpu(device, dev_sc), run_oracle(oracle, dev_sc)
    assert_parity(got, want, name="crates-device-cull")
    assert (got[2].objs.i, got[2].objs.o) == (want[2].objs.i, want[2].objs.o) == (170, len(host_sc.draws))
    host = run_gpu(device, host_sc)
    assert (host[0] == got[0]).all() and host[2].counters() == got[2].counters()
    assert_parity(run_gpu(device, dev_sc, per_draw_sync=True), want, name="crates-device-cull-sync")
    # a bounding box that straddles a plane is not Hidden even when every triangle ends up clipped away
    import dataclasses
    d = dataclasses.replace(scenes.hello_tri().draws[0], bbox=np.array([[-50, -50, -1], [50, 50, 1]], np.float32))
    sc = scenes.Scene("bbox-clipped", 640, 480, rf.FMT_RGBA8888, False, rf.Context(), [d], clear=False)
    got = run_gpu(device, sc)
    assert_parity(got, run_oracle(oracle, sc), name="bbox-clipped")
    assert (got[2].objs.i, got[2].objs.o, int(got[2].calls)) == (1, 1, 1)
    with pytest.raises(rf.RetrofireError) as e:   # the sprite VS has no model-to-projection matrix in u[0..16]
        s = scenes.sprites(10)
        device.render(dataclasses.replace(s.draws[0], bbox=np.zeros((2, 3), np.float32)), device.framebuf(s.w, s.h, s.fmt, True), want_stats=True)
    assert e.value.status == rf.RF_E_INVALID


def test_text_as_textured_geometry(device, oracle):
    """render/text.rs + tex.rs Atlas: the hello.rs demo — glyph quads sampled with SamplerClamp from a font
    atlas, swinging through the frustum (including frames where the text crosses the near plane and is clipped)."""
    for secs in (0.0, 0.7, 2.3, 4.6):
        check(device, oracle, scenes.hello_text(secs))
    big = scenes.hello_text(1.1, msg="\n".join("".join(chr(32 + (r * 7 + c) % 90) for c in range(40)) for r in range(12)))
    check(device, oracle, big)


def test_odd_sized_target(device, oracle):
    """Width/height not multiples of the tile or of 4 (scalar tile I/O path)."""
    sc = scenes.random_soup(500, 333, 211, seed=5, lanes_kind="color3", big=True)
    check(device, oracle, sc)


def test_small_tris_8k_row_bands_cover_the_frame(device, oracle):
    """Sort-first property at full size: rendering the 8 row bands separately and stacking them
    equals the unsharded frame; frags counters add up."""
    from retrofire_b200 import shard
    sc = scenes.small_tris(200_000)
    full_c, full_d, full_s = run_gpu(device, sc)
    acc_c, acc_d = np.zeros_like(full_c), np.zeros_like(full_d)
    fi = fo = 0
    for (y0, y1) in shard.row_bands(sc.h, 8):
        device.set_row_band(y0, y1)
        try:
            c, d, st = run_gpu(device, sc)
        finally:
            device.set_row_band(0, 0xFFFFFFFF)
        acc_c[y0:y1], acc_d[y0:y1] = c[y0:y1], d[y0:y1]
        fi += st.frags.i
        fo += st.frags.o
    assert np.array_equal(acc_c, full_c) and depth_equal(acc_d, full_d)
    assert (fi, fo) == (full_s.frags.i, full_s.frags.o)


def test_empty_and_degenerate(device, oracle):
    """Empty draws, zero-area and all-outside triangles."""
    sc = scenes.hello_tri()
    d = sc.draws[0]
    empty = rf.DrawCall.make(np.zeros((0, 3), np.uint32), d.verts, d.shader, d.uniform[:16].reshape(4, 4), d.viewport)
    degenerate = rf.DrawCall.make([[0, 0, 1], [0, 1, 1], [2, 2, 2]], d.verts, d.shader, d.uniform[:16].reshape(4, 4), d.viewport)
    sc.draws = [empty, degenerate, d]
    check(device, oracle, sc)


def test_index_out_of_bounds_is_an_error(device):
    """render/prim.rs:17-19 panics; the ABI returns RF_E_INDEX_OOB and leaves the target untouched."""
    sc = scenes.hello_tri()
    d = sc.draws[0]
    bad = rf.DrawCall.make([[0, 1, 7]], d.verts, d.shader, d.uniform[:16].reshape(4, 4), d.viewport)
    fb = device.framebuf(sc.w, sc.h, sc.fmt, False)
    with pytest.raises(rf.RetrofireError) as e:
        device.render(bad, fb, want_stats=True)
    assert e.value.status == rf.RF_E_INDEX_OOB
    assert not fb.download_color().any()


def test_target_out_of_bounds_is_an_error(device):
    """A viewport larger than the target makes spans index outside it (render/target.rs:148,173 panics)."""
    sc = scenes.hello_tri()
    d = sc.draws[0]
    fb = device.framebuf(320, 240, sc.fmt, False)
    with pytest.raises(rf.RetrofireError) as e:
        device.render(d, fb, want_stats=True)
    assert e.value.status == rf.RF_E_TARGET_OOB


def test_row_band_sharding_matches_oracle_band(device, oracle):
    """Sort-first sharding: a ctx restricted to a row band renders exactly those rows."""
    sc = scenes.random_soup(1000, 512, 384, seed=41, lanes_kind="color3", big=True)
    device.set_row_band(100, 260)
    try:
        got = run_gpu(device, sc)
    finally:
        device.set_row_band(0, 0xFFFFFFFF)
    want = run_oracle(oracle, sc, band=(100, 260))
    # only the rows of the band are cleared and rasterised by this ctx (the others belong to other ranks)
    band = lambda r: (r[0][100:260], r[1][100:260], r[2])
    assert_parity(band(got), band(want), name="band")
    assert not got[0][:100].any() and not got[0][260:].any(), "rows outside the band must stay untouched"


@pytest.mark.parametrize("dtest", ["less", "none"])
@pytest.mark.parametrize("shift", [40.0, 300.0])
def test_scanlines_above_the_target_are_drawn_at_row_0(device, oracle, dtest, shift):
    """`let y = self.y as usize` (raster.rs:106) saturates: with a viewport that reaches above the target, every scanline with a
    negative y is drawn at row 0, one after the other, in scanline order. Large and small triangles, lit and colour lanes, with
    and without a depth test (without one the last scanline drawn wins row 0); Stats count those fragments like any others."""
    from retrofire_b200 import mathx as mx
    w, h = 160, 96
    for kind in ("color3", "lit"):
        sc0 = scenes.random_soup(400, w, h, seed=21, lanes_kind=kind, big=(kind == "lit"))
        ctx = rf.Context(face_cull=None, depth_test=rf.Ordering.Less if dtest == "less" else None)
        vp = mx.viewport((0, h - shift), (w, -shift))   # the whole picture moved up by `shift` rows
        draws = [dataclasses.replace(d, viewport=np.asarray(vp, dtype=f32).reshape(4, 4), face_cull=0, depth_test=rf.Ordering.Less if dtest == "less" else 0) for d in sc0.draws]
        sc = scenes.Scene(f"above_{kind}_{dtest}_{int(shift)}", w, h, sc0.fmt, True, ctx, draws)
        want = run_oracle(oracle, sc)
        assert want[2].frags.i > 1000
        assert_parity(run_gpu(device, sc), want, name=sc.name)


def test_sampler_once(device, oracle):
    """`SamplerOnce` (tex.rs:313-357, catalogue id RF_FS_TEX_ONCE): unchecked texel fetch. Inside the texture it is a plain lookup; a
    coordinate outside it panics in the reference (slice index) and is RF_E_BAD_TEXTURE on both sides."""
    from retrofire_b200 import mathx as mx
    g = np.random.default_rng(3)
    tex = rf.Texture(g.integers(0, 256, (16, 8, 4), dtype=np.uint8))   # 8 wide, 16 high, Color4 texels
    w, h = 128, 96
    for uvmax, ok in ((0.999, True), (1.25, False)):
        verts = np.array([[-0.9, -0.9, 0.5, 0.0, 0.0], [0.9, -0.9, 0.5, uvmax, 0.0], [0.9, 0.9, 0.5, uvmax, uvmax], [-0.9, 0.9, 0.5, 0.0, uvmax]], dtype=f32)
        tris = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.uint32)
        ctx = rf.Context(face_cull=None)
        shd = rf.shader.new(rf.VS_MVP, rf.FS_TEX_ONCE, texture=tex)
        sc = scenes.Scene(f"once_{uvmax}", w, h, rf.FMT_RGBA8888, True, ctx,
                          [rf.DrawCall.make(tris, verts, shd, np.eye(4, dtype=f32), mx.viewport((0, h), (w, 0)), ctx)])
        if ok:
            want = run_oracle(oracle, sc)
            assert want[2].frags.o > 5000
            assert_parity(run_gpu(device, sc), want, name=sc.name)
        else:
            with pytest.raises(rf.RetrofireError) as eo:
                run_oracle(oracle, sc)
            with pytest.raises(rf.RetrofireError) as eg:
                run_gpu(device, sc)
            assert eo.value.status == eg.value.status == _ffi.RF_E_BAD_TEXTURE
