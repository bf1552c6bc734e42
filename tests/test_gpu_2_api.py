"""GPU parity, part 2: the C-ABI surface around render() — queued passes, rf_render_many, targets of different sizes, page-locked
geometry, strided upload/download, asynchronous downloads, profiling entry points, error statuses."""
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


def test_non_pot_texture_with_repeat_sampler_is_an_error(device):
    """SamplerRepeatPot::new asserts power-of-two dimensions (render/tex.rs:230-231)."""
    sc = scenes.random_soup(10, 64, 64, seed=1, lanes_kind="uv")
    d = sc.draws[0]
    d.shader.texture = rf.Texture(np.zeros((12, 10, 3), np.uint8))
    fb = device.framebuf(64, 64, sc.fmt, True)
    with pytest.raises(rf.RetrofireError) as e:
        device.render(d, fb, want_stats=True)
    assert e.value.status == rf.RF_E_BAD_TEXTURE


def test_unsupported_options_are_reported(device):
    """An out-of-range depth_sort is invalid; a fragment shader with too few lanes is rejected."""
    sc = scenes.hello_tri()
    d = sc.draws[0]
    fb = device.framebuf(sc.w, sc.h, sc.fmt, False)
    import dataclasses
    with pytest.raises(rf.RetrofireError) as e:
        device.render(dataclasses.replace(d, depth_sort=3), fb, want_stats=True)
    assert e.value.status == rf.RF_E_INVALID
    bad = dataclasses.replace(d, shader=rf.shader.new(rf.VS_MVP, rf.FS_TEX_CLAMP_LIT, lanes=3, persp_mask=0))
    with pytest.raises(rf.RetrofireError) as e:
        device.render(bad, fb, want_stats=True)
    assert e.value.status in (rf.RF_E_UNSUPPORTED_SHADER, rf.RF_E_INVALID)


def test_many_small_draws_one_pass(device, oracle):
    """Hundreds of tiny render() calls queued into one pass (the crates demo pattern, crates.rs:114-131)."""
    base = scenes.random_soup(900, 400, 300, seed=77, lanes_kind="lit", big=False)
    d = base.draws[0]
    import dataclasses
    draws = []
    for k in range(300):
        draws.append(dataclasses.replace(d, prims=np.ascontiguousarray(d.prims[3 * k: 3 * k + 3])))
    base.draws = draws
    check(device, oracle, base)


@pytest.mark.parametrize("vary", ["verts", "prims", "both", "neither"])
def test_equal_and_unequal_draw_sizes_in_one_pass(device, oracle, vary):
    """A pass finds a vertex's / primitive's draw by one division when every draw has the same count, by binary search otherwise
    (`find_draw`): all four combinations, with different data per draw."""
    import dataclasses
    draws = []
    for k in range(37):
        nt = 20 + (k % 5 if vary in ("prims", "both") else 0)
        sc = scenes.random_soup(30, 300, 200, seed=500 + k, lanes_kind="color3", big=bool(k & 1))
        d = sc.draws[0]
        verts = d.verts if vary in ("neither", "prims") else np.ascontiguousarray(d.verts[: 90 - 3 * (k % 4)])
        nv = verts.shape[0] // 3
        prims = np.ascontiguousarray(d.prims[np.arange(nt) % nv])          # triangles reused when the draw has fewer than nt
        draws.append(dataclasses.replace(d, prims=prims, verts=verts))
    sc.draws = draws
    check(device, oracle, sc)


def test_render_many_equals_individual_calls(device, oracle):
    """rf_render_many: a frame's list of render() calls in one crossing of the C ABI; re-submitting the same list reuses the
    marshalled array, a changed list does not."""
    sc = scenes.crates("169", 640, 360)
    want = run_oracle(oracle, sc)
    fb = device.framebuf(sc.w, sc.h, sc.fmt, True)
    for rep in range(2):
        fb.clear(sc.ctx)
        device.stats(reset=True)
        device.render_many(sc.draws, fb)
        got = (fb.download_color(), fb.download_depth(), device.stats(reset=True))
        assert_parity(got, want, name=f"render_many-{rep}")
    sc.draws[-1], sc.draws[-2] = sc.draws[-2], sc.draws[-1]       # same list object, same length, different content
    fb.clear(sc.ctx)
    device.render_many(sc.draws, fb)
    assert_parity((fb.download_color(), fb.download_depth(), device.stats(reset=True)), run_oracle(oracle, sc), name="render_many-changed")


def test_targets_of_different_sizes_in_one_pass(device, oracle):
    """Three targets of different sizes and formats drawn in ONE pass: the rasteriser then finds a tile's target by binary
    search over the tile bases (frame batches of equal targets use tile / tiles_per_target instead); twice, so that the second
    pass runs with warm arenas."""
    scs = [scenes.random_soup(400, 320, 200, seed=21, lanes_kind="color3"), scenes.random_soup(300, 96, 50, seed=22, lanes_kind="uv"),
           scenes.random_soup(500, 641, 359, seed=23, lanes_kind="color3", big=True)]
    fbs = [device.framebuf(sc.w, sc.h, sc.fmt, sc.has_depth) for sc in scs]
    try:
        for rep in range(2):
            for sc, fb in zip(scs, fbs):
                fb.clear(sc.ctx)
            device.stats(reset=True)
            for k in range(max(len(sc.draws) for sc in scs)):      # interleaved submission
                for sc, fb in zip(scs, fbs):
                    if k < len(sc.draws):
                        device.render(sc.draws[k], fb)
            total = device.stats(reset=True)
            want_total = rf.Stats()
            for sc, fb in zip(scs, fbs):
                wc, wd, ws = run_oracle(oracle, sc)
                want_total += ws
                assert np.array_equal(fb.download_color(), wc), (sc.name, rep)
                assert depth_equal(fb.download_depth(), wd), (sc.name, rep)
            assert total.counters() == want_total.counters()
    finally:
        for fb in fbs:
            fb._destroy()
            device._targets.remove(fb)


def test_short_uniform_is_zero_padded(device, oracle):
    """A uniform shorter than RF_VS_UNIFORM_F32 floats (one matrix given to the two-matrix solids shader) is zero-padded by
    DrawCall, identically for the device and the oracle: the second matrix is zero, so every normal-derived colour is black."""
    import dataclasses
    mesh = scenes.bunny(subdiv=0, w=400, h=300)
    d = mesh.draws[0]
    one = dataclasses.replace(d, uniform=np.asarray(d.uniform, np.float32).ravel()[:16].reshape(4, 4))
    assert one.uniform.shape == (rf.RF_VS_UNIFORM_F32,) and not one.uniform[16:].any()
    check(device, oracle, dataclasses.replace(mesh, name="bunny-one-matrix", draws=[one]))


def test_page_locked_geometry_is_dmad_directly(device, oracle):
    """Vertex/index arrays in rf_host_alloc memory take the direct-DMA path of rf_render (several draws per pass,
    mixed with pageable draws that go through pinned staging); results are identical."""
    import dataclasses
    a = scenes.random_soup(4000, 640, 360, seed=51, lanes_kind="lit", big=False)
    b = scenes.random_soup(4000, 640, 360, seed=52, lanes_kind="color3", big=True)
    c = scenes.random_soup(4000, 640, 360, seed=53, lanes_kind="lit", big=True)
    want = a
    want.draws = a.draws + b.draws + c.draws

    def pin(x):
        y = device.pinned_empty(x.shape, x.dtype)
        y[...] = x
        return y

    got_scene = dataclasses.replace(want, draws=[dataclasses.replace(want.draws[0], prims=pin(want.draws[0].prims), verts=pin(want.draws[0].verts)),
                                                  want.draws[1],
                                                  dataclasses.replace(want.draws[2], prims=pin(want.draws[2].prims), verts=pin(want.draws[2].verts))])
    assert_parity(run_gpu(device, got_scene), run_oracle(oracle, want), name="direct-dma")
    # opt-in asynchronous form: rf_render does not wait for the DMA; the arrays stay untouched until the sync inside run_gpu
    device.set_geometry_async(True)
    try:
        for _ in range(2):
            assert_parity(run_gpu(device, got_scene), run_oracle(oracle, want), name="direct-dma-async")
    finally:
        device.set_geometry_async(False)


def test_indexed_mesh_paths_screen_vertices_per_vertex(device, oracle):
    """Indexed meshes (>= 2 uses per vertex) take the k_vertex -> k_assemble<LT, true> path (to_screen once per vertex):
    alone, depth-sorted without a depth test (Render::depth of unclipped triangles reads clip-space z), pushed through the
    near plane so that part of the mesh is clipped, with a bounding box, and mixed with a vertex-per-triangle soup in one
    pass (which switches the whole pass back to per-primitive to_screen)."""
    import dataclasses
    mesh = scenes.bunny(subdiv=0, w=800, h=600)
    check(device, oracle, mesh)
    d = mesh.draws[0]
    nod = rf.Context(depth_sort=rf.DepthSort.BackToFront, depth_test=None, face_cull=None)
    check(device, oracle, dataclasses.replace(mesh, name="bunny-sorted", ctx=nod, draws=[dataclasses.replace(
        d, depth_sort=int(rf.DepthSort.BackToFront), depth_test=0, face_cull=0)]))
    from retrofire_b200 import mathx as mx
    m0 = np.asarray(d.uniform, np.float32).ravel()[:16].reshape(4, 4)
    for tz in (2.0, 2.4):    # towards the camera: 2.0 crosses the side planes, 2.4 also the near plane (129 vertices behind it)
        near = dataclasses.replace(d, uniform=mx.then(mx.translate3(0.0, 0.0, tz), m0))
        got = run_gpu(device, dataclasses.replace(mesh, name="bunny-near", draws=[near]))
        assert_parity(got, run_oracle(oracle, dataclasses.replace(mesh, draws=[near])), name=f"bunny-near-{tz}")
        assert got[2].frags.i > 50000
    lo, hi = d.verts[:, :3].min(0), d.verts[:, :3].max(0)
    check(device, oracle, dataclasses.replace(mesh, name="bunny-bbox", draws=[dataclasses.replace(d, bbox=np.stack([lo, hi]))]))
    soup = scenes.random_soup(1500, 800, 600, seed=3, lanes_kind="color3", big=True)
    check(device, oracle, dataclasses.replace(mesh, name="bunny+soup", draws=[d] + soup.draws + [d]))


@pytest.mark.parametrize("fmt", [rf.FMT_RGBA8888, rf.FMT_XRGB8888, rf.FMT_RGB888, rf.FMT_RGB565])
def test_strided_upload_render_download(device, oracle, fmt):
    """A frontend-owned pixel slice with a row stride larger than the width (front/src/sdl2.rs:208-217, util/buf.rs:437-439):
    existing colour and depth contents are uploaded with a stride, a frame is rendered over them WITHOUT a clear (the uploaded
    depth blocks part of it), and colour / depth are downloaded into strided buffers whose padding must stay untouched."""
    from oracle import rfo
    w, h, cs, ds = 150, 90, 157, 153
    g = np.random.default_rng(int(fmt) + 5)
    sc = scenes.random_soup(600, w, h, seed=31, lanes_kind="color3", big=True)
    sc.fmt, sc.clear = fmt, False
    cont = g.integers(0, 1 << 32, (h, w), dtype=np.uint64).astype(np.uint32) & np.uint32(0xFFFF if fmt == rf.FMT_RGB565 else 0xFFFFFF)
    depth0 = np.where(g.integers(0, 3, (h, w)) == 0, np.float32(np.inf), g.uniform(0, 0.2, (h, w)).astype(f32)).astype(f32)
    # oracle: a host target that starts with these contents
    tgt = oracle.HostTarget(w, h, fmt, True)
    tgt.color[:] = cont; tgt.depth[:] = depth0
    want_stats = rf.Stats()
    for d in sc.draws:
        want_stats += oracle.render(d, tgt)
    # device: strided upload, render, strided download
    host0 = rfo.container_to_host(fmt, cont)
    pad_shape = (h, cs) + host0.shape[2:]
    hostbuf = np.full(pad_shape, 0xA5, dtype=host0.dtype); hostbuf[:, :w] = host0
    depthbuf = np.full((h, ds), -7.0, f32); depthbuf[:, :w] = depth0
    fb = device.framebuf(w, h, fmt, True)
    try:
        device._check(device.lib.rf_target_upload_color(device.h, fb.h, hostbuf.ctypes.data, cs))
        device._check(device.lib.rf_target_upload_depth(device.h, fb.h, depthbuf.ctypes.data, ds))
        device.stats(reset=True)
        for d in sc.draws:
            device.render(d, fb)
        got_stats = device.stats(reset=True)
        out = np.full(pad_shape, 0x5A, dtype=host0.dtype); dout = np.full((h, ds), -9.0, f32)
        device._check(device.lib.rf_target_download_color(device.h, fb.h, out.ctypes.data, cs))
        device._check(device.lib.rf_target_download_depth(device.h, fb.h, dout.ctypes.data, ds))
    finally:
        fb._destroy(); device._targets.remove(fb)
    assert np.array_equal(out[:, :w], tgt.host_color()) and depth_equal(dout[:, :w], tgt.depth)
    assert (out[:, w:] == 0x5A).all() and (dout[:, w:] == -9.0).all(), "row padding must not be written"
    assert got_stats.counters() == want_stats.counters() and 0 < want_stats.frags.o < want_stats.frags.i


def test_async_download_profiling_and_device_pointers(device, oracle):
    """The entry points bench.py's end-to-end and per-kernel legs rely on: `rf_target_download_color_async` into page-locked
    memory (valid after `rf_sync`), `rf_ctx_profile` / `rf_ctx_kernel_times` / `rf_kernel_name`, `rf_ctx_last_pass`, and the raw
    device pointers of a target. Profiling serialises the pass; the frame must not change."""
    sc = scenes.random_soup(800, 320, 200, seed=12, lanes_kind="uv", big=True)
    want = run_oracle(oracle, sc)
    fb = device.framebuf(sc.w, sc.h, sc.fmt, True)
    try:
        for level in (2, 1, 0):
            device.profile(level)
            device.kernel_times()                                   # reset the accumulators
            fb.clear(sc.ctx)
            device.stats(reset=True)
            for d in sc.draws:
                device.render(d, fb)
            out = device.pinned_empty((sc.h, sc.w, 4), np.uint8)
            out[:] = 0x5A
            fb.download_color_async(out)
            device.sync()
            stats = device.stats(reset=True)
            assert np.array_equal(out, want[0]) and np.array_equal(fb.download_color(), want[0]), level
            assert depth_equal(fb.download_depth(), want[1]), level
            assert stats.counters() == want[2].counters(), level
            times = device.kernel_times()
            assert len(times) == rf._ffi.RF_N_KERNELS and "k_raster" in times and "k_setup" in times
            if level:
                assert times["k_raster"][1] >= 1, times
            if level == 2:
                assert all(times[k][1] >= 1 for k in ("k_vertex", "k_assemble", "k_setup")), times
            if level == 0:
                assert all(n == 0 for _, n in times.values()), times
            ns, launches = device.last_pass()
            assert launches >= 5
        cp, dp = fb.color_devptr(), fb.depth_devptr()
        assert cp and dp and cp != dp
    finally:
        device.profile(0)
        fb._destroy(); device._targets.remove(fb)


def test_two_clears_without_a_draw_keep_the_last_value(device):
    """Frame::clear twice (front/src/lib.rs:103-120): the second value stays. Both clears are recorded in one pass, whose clears run
    in a single launch — the later one must replace the earlier one, not race it."""
    fb = device.framebuf(200, 100, rf.FMT_RGBA8888, True)
    for rep in range(3):
        fb.clear(rf.Context(color_clear=(10, 20, 30, 40), depth_clear=2.0))
        fb.clear(rf.Context(color_clear=(200, 100, 50, 255), depth_clear=None))     # colour only: the depth of the first clear stays
        fb.clear(rf.Context(color_clear=None, depth_clear=4.0))                     # depth only
        c, d = fb.download_color(), fb.download_depth()
        assert (c == np.array([200, 100, 50, 255], np.uint8)).all()
        assert (d == np.float32(0.25)).all()
        fb.clear(rf.Context(color_clear=(1, 2, 3, 4), depth_clear=8.0))
    c, d = fb.download_color(), fb.download_depth()
    assert (c == np.array([1, 2, 3, 4], np.uint8)).all() and (d == np.float32(0.125)).all()


def test_lazy_depth_tiles_through_later_passes(device, oracle):
    """Lazy depth clear: a clear recorded with draws marks the tiles the pass does not touch instead of filling their
    depth. Every way such a tile is used afterwards must see the clear value: a later pass without a clear (one warp per tile, and a
    heaviest tile cut into row slices), a download, a second clear with another value, a clear without draws (the plane is written
    as memory), an upload over marked tiles, and the raw device pointer (which ends laziness for the target). One device target
    and one oracle target go through the same sequence; colour, depth and Stats are compared after every step."""
    w, h = 256, 192
    few = scenes.random_soup(40, w, h, seed=61, lanes_kind="color3", big=False)       # most of the 48 tiles stay untouched
    some = scenes.random_soup(300, w, h, seed=63, lanes_kind="uv", big=True)
    many = scenes.random_soup(6000, w, h, seed=62, lanes_kind="color3", big=True)     # >= 160 triangles per tile: row-slice tasks
    fb = device.framebuf(w, h, few.fmt, True)
    ref = oracle.HostTarget(w, h, few.fmt, True)
    dstat, rstat = rf.Stats(), rf.Stats()

    def draw(sc):
        nonlocal rstat
        for d in sc.draws:
            device.render(d, fb)
            rstat += oracle.render(d, ref)

    def clear(color, depth):
        fb.clear(rf.Context(color_clear=color, depth_clear=depth))
        ref.clear(color, depth)

    def same(step, depth=True):
        nonlocal dstat
        dstat += device.stats(reset=True)
        assert np.array_equal(fb.download_color(), ref.host_color()), step
        if depth:
            assert depth_equal(fb.download_depth(), ref.depth), step
        assert dstat.counters() == rstat.counters(), step

    try:
        device.stats(reset=True)
        clear((10, 20, 30, 255), 8.0); draw(few); device.flush()
        draw(some); device.flush()                                   # no clear: marked tiles are first touched here
        same("later pass, one warp per tile")
        clear(None, 4.0); draw(few); device.flush()                  # marked again, with another value
        draw(many); device.flush()                                   # no clear: marked tiles rasterised as row slices
        same("later pass, row slices", depth=False)                  # colour first: the depth download below materialises
        draw(few)
        same("after the slices")
        clear((1, 2, 3, 4), 2.0); draw(few); device.flush()          # marked
        clear((5, 6, 7, 8), 16.0); device.flush()                    # a clear without draws writes the whole plane
        draw(some)
        same("clear without draws over marked tiles")
        clear(None, 2.0); draw(few); device.flush()                  # marked
        up = np.random.default_rng(7).uniform(0.05, 0.6, (h, w)).astype(f32)
        fb.upload_depth(up); ref.depth[:] = up
        draw(some)
        same("upload over marked tiles")
        clear((9, 9, 9, 255), 32.0); draw(few); device.flush()       # marked
        assert fb.depth_devptr()                                      # handed out: written out, and kept as memory from here on
        draw(some); device.flush()
        clear(None, 8.0); draw(few); device.flush(); draw(some)
        same("after the device pointer was handed out")
    finally:
        fb._destroy(); device._targets.remove(fb)


def test_batch_builder_mirrors_render(device, oracle):
    """`Batch` (batch.rs:31-147): the builder's render() is render() with the same arguments; clone() leaves the original usable
    (crates.rs:103-130 clones one batch per crate). Stats accumulate in the Context like render.rs:206."""
    sc = scenes.random_soup(500, 320, 200, seed=11, lanes_kind="color3", big=False)
    d = sc.draws[0]
    want = run_oracle(oracle, sc)
    fb = device.framebuf(sc.w, sc.h, sc.fmt, sc.has_depth)
    fb.clear(sc.ctx)
    ctx = rf.Context(face_cull=d.face_cull or None, depth_test=d.depth_test or None)
    base = rf.Batch().mesh((d.prims, d.verts)).shader(d.shader).viewport(d.viewport).context(ctx)
    b = base.clone().uniform(d.uniform).target(fb)
    b.render()
    assert base.target_ is None and base.uniform_ is None, "builder steps return new batches"
    assert np.array_equal(fb.download_color(), want[0]) and depth_equal(fb.download_depth(), want[1])
    assert ctx.stats.counters() == want[2].counters()
    # primitives() / vertices() instead of mesh(), queued (no sync), twice: Stats double
    fb.clear(sc.ctx)
    device.stats(reset=True)
    b2 = rf.Batch().primitives(d.prims).vertices(d.verts).shader(d.shader).viewport(d.viewport).uniform(d.uniform).target(fb).context(ctx)
    b2.render(sync_stats=False)
    b2.render(sync_stats=False)
    got = device.stats(reset=True)
    assert got.prims.i == 2 * want[2].prims.i and got.frags.i == 2 * want[2].frags.i
    assert np.array_equal(fb.download_color(), want[0]) and depth_equal(fb.download_depth(), want[1])
    fb._destroy(); device._targets.remove(fb)


def test_resident_mesh_remembers_its_primitive_kind(device, oracle):
    """rf_mesh_create records RF_PRIM_TRIS / RF_PRIM_EDGES (ABI 2): an Edge mesh draws like host-pointer edges, and naming a mesh
    with the other rf_draw.prim_kind is RF_E_INVALID instead of reinterpreting the index list."""
    sc = scenes.random_lines(300, 320, 200, seed=5, lanes_kind="color3")
    d = sc.draws[0]
    want = run_oracle(oracle, sc)
    m = device.mesh(d.prims, d.verts, edges=True)
    res = dataclasses.replace(d, mesh=m)
    got = run_gpu(device, dataclasses.replace(sc, draws=[res]))
    assert_parity(got, want, name="edge mesh")
    fb = device.framebuf(sc.w, sc.h, sc.fmt, sc.has_depth)
    with pytest.raises(rf.RetrofireError) as e:
        device.render(dataclasses.replace(res, prim_kind=_ffi.PRIM_TRIS), fb)
    assert e.value.status == _ffi.RF_E_INVALID
    tri = scenes.random_soup(100, 320, 200, seed=3, lanes_kind="color3", big=False).draws[0]
    mt = device.mesh(tri.prims, tri.verts)
    with pytest.raises(rf.RetrofireError) as e:
        device.render(dataclasses.replace(tri, mesh=mt, prim_kind=_ffi.PRIM_EDGES), fb)
    assert e.value.status == _ffi.RF_E_INVALID
    device.sync()
    fb._destroy(); device._targets.remove(fb)


def test_closed_device_handles_are_not_reused(oracle):
    """Texture / DrawCall caches are keyed on the Device instance, not on id(): a second Device created after the first was
    closed must upload its own texture and marshal its own rf_draw (use-after-free through a recycled id)."""
    sc = scenes.textured_quad()
    want = run_oracle(oracle, sc)
    for _ in range(3):
        with rf.Device(0) as dev:
            assert_parity(run_gpu(dev, sc), want, name=sc.name)


def test_overflowing_pass_after_a_small_one_in_a_fresh_context(oracle):
    """The flow that exposed the k_setup poison race (profiles/r02_memcheck.txt): a context whose arenas were sized by a small pass of
    3-lane triangles gets a pass of four 4K crates frames (5 lanes, 500,000 span records) — k_setup overflows the arenas while its own
    blocks are still starting, the pass is replayed several times with larger arenas. Every frame must equal the oracle's, the
    arena requests must stay sane (the race produced requests of hundreds of GB -> RF_E_NOMEM) and nothing may be written out of
    bounds (it also produced triangle records far outside the arena: a sticky launch failure). Three fresh contexts."""
    small = scenes.bunny(subdiv=0, w=400, h=300)
    big = scenes.crates("1089")
    want_small, want_big = run_oracle(oracle, small), run_oracle(oracle, big)
    for rep in range(3):
        with rf.Device(0) as dev:
            assert_parity(run_gpu(dev, small), want_small, name=f"{small.name}-{rep}")
            fbs = [dev.framebuf(big.w, big.h, big.fmt, True) for _ in range(4)]
            try:
                dev.stats(reset=True)
                for fb in fbs:
                    fb.clear(big.ctx)
                for fb in fbs:
                    dev.render_many(big.draws, fb)
                dev.flush()
                dev.sync()
                assert dev.replays() >= 1, "the second pass was meant to overflow the arenas of the first"
                stats = dev.stats(reset=True)
                assert stats.frags.i == 4 * want_big[2].frags.i and stats.frags.o == 4 * want_big[2].frags.o
                for k, fb in enumerate(fbs):
                    got = (fb.download_color(), fb.download_depth(), want_big[2])
                    assert_parity(got, want_big, name=f"{big.name}-{rep}-{k}")
            finally:
                for fb in fbs:
                    fb._destroy(); dev._targets.remove(fb)
