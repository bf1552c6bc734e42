"""GPU parity, part 3 (collected last): adversarial inputs — half-pixel lattices, ties on every strict comparison, NaN/inf vertices,
extreme aspect ratios, very deep tile bins, arena growth and replay, the seeded fuzzers."""
import os

import numpy as np
import pytest

import retrofire_b200 as rf
from retrofire_b200 import scenes
from tests.parity import assert_parity, depth_equal, run_gpu, run_oracle

f32 = np.float32

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def check(device, oracle, sc, **kw):
    assert_parity(run_gpu(device, sc), run_oracle(oracle, sc), name=sc.name, **kw)


def test_depth_sort_mixed_pass_and_ties(device, oracle):
    """One pass holding a back-to-front draw, an unsorted draw, a line draw (edges all have depth +inf and keep their
    order) and a front-to-back draw, no depth test; then a stack of coplanar overlapping quads whose triangles all
    have EQUAL depth: ties keep primitive order (the reference's sort_unstable_by leaves them unspecified)."""
    import dataclasses
    nod = dict(depth_test=None, face_cull=None)
    a = scenes.random_soup(700, 512, 384, seed=41, lanes_kind="lit", big=True, ctx=rf.Context(depth_sort=rf.DepthSort.BackToFront, **nod))
    b = scenes.random_soup(700, 512, 384, seed=42, lanes_kind="color3", big=True, ctx=rf.Context(**nod))
    l = scenes.random_lines(600, 512, 384, seed=43, ctx=rf.Context(depth_sort=rf.DepthSort.BackToFront, **nod))
    c = scenes.random_soup(700, 512, 384, seed=44, lanes_kind="disc", big=True, ctx=rf.Context(depth_sort=rf.DepthSort.FrontToBack, **nod))
    a.draws += b.draws + l.draws + c.draws
    a.name = "depth-sort-mixed"
    want = run_oracle(oracle, a)
    assert_parity(run_gpu(device, a), want, name="mixed-queued")
    assert_parity(run_gpu(device, a, per_draw_sync=True), want, name="mixed-sync")

    g = np.random.default_rng(9)
    n = 200
    quads, prims = [], []
    for k in range(n):
        cx, cy, r = g.uniform(-0.6, 0.6), g.uniform(-0.6, 0.6), g.uniform(0.1, 0.4)
        col = g.uniform(0, 1, 3)
        for dx, dy in ((-r, -r), (r, -r), (r, r), (-r, r)):
            quads.append([cx + dx, cy + dy, 0.25, *col])
        prims += [[4 * k, 4 * k + 1, 4 * k + 2], [4 * k, 4 * k + 2, 4 * k + 3]]
    from retrofire_b200 import mathx as mx
    for order in (rf.DepthSort.BackToFront, rf.DepthSort.FrontToBack):
        ctx = rf.Context(depth_sort=order, **nod)
        call = rf.DrawCall.make(np.array(prims, np.uint32), np.array(quads, np.float32), rf.shader.new(rf.VS_MVP, rf.FS_COLOR3F),
                                mx.identity(), mx.viewport((0, 0), (256, 256)), ctx)
        check(device, oracle, scenes.Scene("depth-sort-ties", 256, 256, rf.FMT_RGBA8888, True, ctx, [call]))


def test_frame_batch_with_empty_and_nan_frames(device, oracle):
    """A frame batch in which some frames draw nothing (the mesh is behind the camera, or scaled to a point) and one frame's
    matrix is NaN: every frame must still equal its own oracle frame, and the empty ones must keep their clear values."""
    import dataclasses
    verts, faces = scenes.bunny_mesh(0)
    base = scenes.bunny(subdiv=0, theta=0.3, w=320, h=200)
    d0 = base.draws[0]
    uniforms = []
    for f in range(9):
        u = scenes.bunny(subdiv=0, theta=0.4 * f, w=320, h=200).draws[0].uniform.copy()
        if f in (1, 5):
            u[:16] = -u[:16]            # clip position negated: w < 0, everything outside
        if f == 3:
            u[:16] = 0; u[15] = 1       # every vertex at the clip-space origin: zero-area triangles
        if f == 7:
            u[:16] = np.nan
        uniforms.append(u)
    mesh = device.mesh(faces, verts)
    call = dataclasses.replace(d0, mesh=mesh)
    targets = [device.framebuf(320, 200, base.fmt, True) for _ in uniforms]
    try:
        for t in targets:
            t.clear(base.ctx)
        device.stats(reset=True)
        device.render_frames(call, targets, np.stack(uniforms))
        got_stats = device.stats(reset=True)
        want_total = rf.Stats()
        for f, (u, t) in enumerate(zip(uniforms, targets)):
            sc = scenes.Scene(f"frame-{f}", 320, 200, base.fmt, True, base.ctx, [dataclasses.replace(d0, uniform=u)])
            wc, wd, ws = run_oracle(oracle, sc)
            want_total += ws
            assert np.array_equal(t.download_color(), wc), f
            assert depth_equal(t.download_depth(), wd), f
        assert got_stats.counters() == want_total.counters()
    finally:
        for t in targets:
            t._destroy(); device._targets.remove(t)


def test_every_tile_heaviest_and_repeated_passes(device, oracle):
    """Big overlapping triangles put >= 160 triangles into nearly every tile, so nearly every tile is split into row-slice
    tasks (regression: the task list used to hold one word per tile and overflowed, corrupting later passes). The same frame
    is rendered several times through the same context: every repetition must equal the oracle."""
    a = scenes.random_soup(4000, 640, 360, seed=51, lanes_kind="lit", big=False)
    b = scenes.random_soup(4000, 640, 360, seed=52, lanes_kind="color3", big=True)
    c = scenes.random_soup(4000, 640, 360, seed=53, lanes_kind="lit", big=True)
    a.draws = a.draws + b.draws + c.draws
    want = run_oracle(oracle, a)
    for rep in range(4):
        assert_parity(run_gpu(device, a), want, name=f"heaviest-everywhere-{rep}")
    small = scenes.random_soup(6000, 96, 64, seed=54, lanes_kind="color3", big=True)     # 6 tiles, all of them heaviest
    want = run_oracle(oracle, small)
    for rep in range(3):
        assert_parity(run_gpu(device, small), want, name=f"six-heaviest-tiles-{rep}")


def test_very_deep_tile_bin(device, oracle):
    """40,000 triangles stacked on the same few tiles: bins deeper than one shared-memory sort run (16,384)
    are merged from sorted runs; submission order must still be exact (depth ties, frags.o)."""
    from retrofire_b200 import mathx as mx
    g = np.random.default_rng(5)
    n, w, h = 40000, 256, 192
    z = g.choice(np.array([2.0, 3.0, 4.0], dtype=np.float32), (n, 1))          # many exact depth ties
    c = g.uniform(-0.05, 0.05, (n, 1, 2)).astype(np.float32) * z[:, :, None]
    xy = c + g.uniform(-0.04, 0.04, (n, 3, 2)).astype(np.float32) * z[:, :, None]
    pos = np.concatenate([xy, np.repeat(z[:, :, None], 3, 1)], 2)
    col = g.uniform(0, 1, (n, 3, 3)).astype(np.float32)
    verts = np.concatenate([pos, col], 2).reshape(3 * n, 6).astype(np.float32)
    tris = np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
    ctx = rf.Context(face_cull=None)
    sc = scenes.Scene("deep_bin", w, h, rf.FMT_RGBA8888, True, ctx,
                      [rf.DrawCall.make(tris, verts, rf.shader.new(rf.VS_MVP, rf.FS_COLOR3F), mx.perspective(1.0, w / h, 0.5, 50.0),
                                        mx.viewport((0, h), (w, 0)), ctx)])
    check(device, oracle, sc)


def lattice_scene(size, persp, dtest):
    """Every tie the fill rule has to break, thousands of times: vertices on a half-pixel lattice (pixel centres, pixel corners,
    horizontal and vertical edges through centres), coordinates exactly on and just outside the six clip planes (on-plane =
    inside, clip.rs:108-111), zero-area and repeated triangles (equal depths: `curr < new` fails the second one, ctx.rs:86-89),
    w = 0 vertices in the perspective variant. Power-of-two targets make the viewport arithmetic exact; the 33x17 one does not.
    Without a depth test every fragment is written, so the frame is decided by submission order alone, ≈ 100 layers deep."""
    w, h = size
    g = np.random.default_rng(w * 1000 + h + int(persp))
    n = 3000
    lx = g.integers(-w - 4, w + 5, (n, 3, 1)).astype(f32) / f32(w)        # screen x = k/2 pixels, a few columns beyond the planes
    ly = g.integers(-h - 4, h + 5, (n, 3, 1)).astype(f32) / f32(h)
    c = g.integers(0, 4, (n, 1, 2)).astype(f32) / f32(4)
    small = g.integers(0, 2, (n, 1, 1)).astype(bool)                       # half of them small: centre + a few lattice steps
    sx = (g.integers(-w, w, (n, 1, 1)) + g.integers(-6, 7, (n, 3, 1))).astype(f32) / f32(w)
    sy = (g.integers(-h, h, (n, 1, 1)) + g.integers(-6, 7, (n, 3, 1))).astype(f32) / f32(h)
    x = np.where(small, sx, lx); y = np.where(small, sy, ly)
    z = g.choice(np.array([-1.25, -1.0, -0.5, 0.0, 0.25, 0.5, 1.0, 1.25], f32), (n, 3, 1))
    z = np.where(g.integers(0, 3, (n, 1, 1)) == 0, z[:, :1], z)            # a third of them at constant depth
    pos = np.concatenate([x, y, z], 2).astype(f32)
    pos[n // 2:n // 2 + 200] = pos[:200]                                   # exact repeats later in submission order
    pos[100:140, 2] = pos[100:140, 0]                                      # zero-area: two equal vertices
    attr = g.integers(0, 5, (n, 3, 3)).astype(f32) / f32(4)
    verts = np.concatenate([pos, attr], 2).reshape(3 * n, -1).astype(f32)
    tris = np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
    mvp = np.eye(4, dtype=f32)
    if persp:
        mvp[3] = [0, 0, 1, 1]                                              # w = z + 1: 0 at z = -1, the near plane at z = -0.5
    # depth-sorted variants: z comes from eight lattice values, so hundreds of triangles share a sort key and only the defined
    # tie order (primitive order, DESIGN §4 "Depth sort") keeps the frame deterministic
    dsort = {"none-back-to-front": rf.DepthSort.BackToFront, "none-front-to-back": rf.DepthSort.FrontToBack}.get(dtest)
    ctx = rf.Context(face_cull=None, depth_test=rf.Ordering.Less if dtest == "less" else None, depth_sort=dsort)
    shd = rf.shader.new(rf.VS_MVP, rf.FS_COLOR3F)
    from retrofire_b200 import mathx as mx
    sc = scenes.Scene(f"lattice_{w}x{h}_{int(persp)}", w, h, rf.FMT_RGBA8888, True, ctx,
                      [rf.DrawCall.make(tris, verts, shd, mvp, mx.viewport((0, h), (w, 0)), ctx)])
    return sc


@pytest.mark.parametrize("dtest", ["less", "none", "none-back-to-front", "none-front-to-back"])
@pytest.mark.parametrize("persp", [False, True])
@pytest.mark.parametrize("size", [(64, 32), (128, 128), (33, 17)])
def test_lattice_ties(device, oracle, size, persp, dtest):
    """See lattice_scene."""
    w, h = size
    sc = lattice_scene(size, persp, dtest)
    try:
        want = run_oracle(oracle, sc)
    except rf.RetrofireError as e:      # the reference would panic (span outside the target): the ABI must report the same
        with pytest.raises(rf.RetrofireError) as ge:
            run_gpu(device, sc)
        assert ge.value.status == e.status
        return
    assert want[2].frags.i > min(1000, w * h * 50), "the lattice scene must actually rasterise"
    assert_parity(run_gpu(device, sc), want, name=sc.name)


@pytest.mark.parametrize("size", [(1, 1), (2, 3), (7, 5), (31, 33), (32, 32), (65, 1), (1, 70)])
def test_lattice_ties_tiny_targets(device, oracle, size):
    """Targets smaller than, equal to and one pixel past a 32x32 tile, and one-pixel-wide strips."""
    test_lattice_ties(device, oracle, size, True, "less")
    test_lattice_ties(device, oracle, size, False, "none")


@pytest.mark.parametrize("persp", [False, True])
@pytest.mark.parametrize("kind", ["checker", "disc", "uv_repeat", "texclamp", "lit"])
def test_lattice_shader_ties(device, oracle, kind, persp):
    """Shader decisions on exact ties: attribute values on a quarter lattice over half-pixel lattice triangles, so interpolated
    values land exactly on the checker's 0.5 (crates.rs:33-36, strict `>`), on the sprite disc's d2 = 1 (sprites.rs:46-52, strict
    `<` or discard), on texel boundaries of the clamp / repeat samplers (tex.rs:218-304) and on negative coordinates."""
    w, h = 64, 64
    g = np.random.default_rng(len(kind) * 10 + int(persp))
    n = 1500
    cx = g.integers(-w, w, (n, 1, 1)); cy = g.integers(-h, h, (n, 1, 1))
    x = (cx + g.integers(-16, 17, (n, 3, 1))).astype(f32) / f32(w)
    y = (cy + g.integers(-16, 17, (n, 3, 1))).astype(f32) / f32(h)
    z = g.choice(np.array([-0.5, 0.0, 0.0, 0.25, 0.5, 1.0], f32), (n, 3, 1))
    z = np.where(g.integers(0, 2, (n, 1, 1)) == 0, z[:, :1], z)
    uv = g.integers(-4, 9, (n, 3, 2)).astype(f32) / f32(4)                  # -1 .. 2 in quarters
    if kind == "checker":
        attr, shd = uv, rf.shader.new(rf.VS_MVP, rf.FS_CHECKER)
    elif kind == "disc":
        attr, shd = g.integers(-4, 5, (n, 3, 2)).astype(f32) / f32(4), rf.shader.new(rf.VS_MVP, rf.FS_SPRITE_DISC)
    elif kind == "uv_repeat":
        attr, shd = uv, rf.shader.new(rf.VS_MVP, rf.FS_TEX_REPEAT_POT, texture=rf.Texture(g.integers(0, 256, (8, 8, 3), dtype=np.uint8)))
    elif kind == "texclamp":
        attr, shd = uv, rf.shader.new(rf.VS_MVP, rf.FS_TEX_CLAMP, texture=rf.Texture(g.integers(0, 256, (8, 4, 4), dtype=np.uint8)))
    else:
        nrm = g.integers(-2, 3, (n, 3, 3)).astype(f32) / f32(2)
        attr = np.concatenate([nrm, uv], 2)
        shd = rf.shader.new(rf.VS_MVP, rf.FS_TEX_CLAMP_LIT, fs_uniform=[0.0, 0.0, -1.0], texture=rf.Texture(g.integers(0, 256, (16, 16, 3), dtype=np.uint8)))
    verts = np.concatenate([x, y, z, attr.astype(f32)], 2).reshape(3 * n, -1).astype(f32)
    tris = np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
    mvp = np.eye(4, dtype=f32)
    if persp:
        mvp[3] = [0, 0, 1, 1]
    from retrofire_b200 import mathx as mx
    ctx = rf.Context(face_cull=None, depth_test=rf.Ordering.Less if persp else None)   # w = 1 everywhere: equal depths, shade them all
    sc = scenes.Scene(f"lattice_{kind}_{int(persp)}", w, h, rf.FMT_RGBA8888, True, ctx,
                      [rf.DrawCall.make(tris, verts, shd, mvp, mx.viewport((0, h), (w, 0)), ctx)])
    try:
        want = run_oracle(oracle, sc)
    except rf.RetrofireError as e:
        with pytest.raises(rf.RetrofireError) as ge:
            run_gpu(device, sc)
        assert ge.value.status == e.status
        return
    assert want[2].frags.i > 1000
    assert_parity(run_gpu(device, sc), want, name=sc.name)


@pytest.mark.parametrize("persp", [False, True])
@pytest.mark.parametrize("size", [(64, 32), (33, 17)])
def test_lattice_lines(device, oracle, size, persp):
    """Line segments (raster.rs:122-177) with end points on the half-pixel lattice: exact 45-degree slopes (|dx| == |dy| picks the
    major axis by a tie), axis-aligned and zero-length segments, end points on and beyond the near/far planes; mixed with lattice
    triangles submitted before them, no depth test, so submission order decides every pixel."""
    w, h = size
    g = np.random.default_rng(w * 77 + h + int(persp))
    n = 2500
    # end points stay inside the x/y planes (a segment ending on the right or bottom plane indexes one past the target and panics
    # in the reference); z crosses the near and far planes
    x0 = g.integers(-w + 12, w - 11, (n, 1, 1)); y0 = g.integers(-h + 12, h - 11, (n, 1, 1))
    dx = g.integers(-10, 11, (n, 1, 1)); dy = g.integers(-10, 11, (n, 1, 1))
    k = n // 5
    dy[:k] = dx[:k]; dy[k:2 * k] = -dx[k:2 * k]; dy[2 * k:2 * k + 100] = 0; dx[2 * k + 100:2 * k + 200] = 0; dx[2 * k + 200:2 * k + 220] = 0; dy[2 * k + 200:2 * k + 220] = 0
    x = np.concatenate([x0, x0 + dx], 1).astype(f32) / f32(w)
    y = np.concatenate([y0, y0 + dy], 1).astype(f32) / f32(h)
    z = g.choice(np.array([-1.25, -1.0, -0.5, 0.0, 0.25, 0.5, 1.0, 1.25], f32), (n, 2, 1))
    attr = g.integers(0, 5, (n, 2, 3)).astype(f32) / f32(4)
    verts = np.concatenate([x, y, z, attr], 2).reshape(2 * n, -1).astype(f32)
    edges = np.arange(2 * n, dtype=np.uint32).reshape(n, 2)
    g.shuffle(edges)
    mvp = np.eye(4, dtype=f32)
    if persp:
        mvp[3] = [0, 0, 1, 1]
    from retrofire_b200 import mathx as mx
    ctx = rf.Context(face_cull=None, depth_test=None)
    shd = rf.shader.new(rf.VS_MVP, rf.FS_COLOR3F)
    vp = mx.viewport((0, h), (w, 0))
    tri_verts = verts[: 3 * (2 * n // 3)]
    tris = np.arange(tri_verts.shape[0], dtype=np.uint32).reshape(-1, 3)[:300]
    sc = scenes.Scene(f"lattice_lines_{w}x{h}_{int(persp)}", w, h, rf.FMT_RGBA8888, True, ctx,
                      [rf.DrawCall.make(tris, tri_verts, shd, mvp, vp, ctx), rf.DrawCall.make(edges, verts, shd, mvp, vp, ctx, edges=True)])
    try:
        want = run_oracle(oracle, sc)
    except rf.RetrofireError as e:
        with pytest.raises(rf.RetrofireError) as ge:
            run_gpu(device, sc)
        assert ge.value.status == e.status
        return
    assert want[2].frags.i > 1000
    assert_parity(run_gpu(device, sc), want, name=sc.name)


@pytest.mark.parametrize("persp", [False, True])
@pytest.mark.parametrize("size", [(64, 64), (200, 120)])
def test_lattice_grid_mesh_shared_edges(device, oracle, size, persp):
    """An indexed grid mesh (every interior vertex shared by six triangles: the screen-vertex path of k_assemble) whose vertices are
    jittered on the half-pixel lattice and whose border lies exactly on the clip planes, no depth test. The reference's own
    "no gaps, no overdraw" KAT (raster.rs:326-368) does NOT generalise to such a mesh: a shared edge is the long edge of one
    triangle (split at `mid1`, a rounded point, and re-walked from there) and a short edge of its neighbour, so where it passes
    exactly through pixel centres the two walks can disagree — the oracle covers 4,110 fragments on the 4,096-pixel target. The
    device path has to reproduce exactly that: same double-covered pixels, same winner by submission order, same counters."""
    w, h = size
    cell = 4                                              # pixels per cell
    nx, ny = w // cell, h // cell
    g = np.random.default_rng(w + h + int(persp))
    gx, gy = np.meshgrid(np.arange(nx + 1) * cell * 2, np.arange(ny + 1) * cell * 2)    # in half pixels
    jx = g.integers(-2, 3, gx.shape); jy = g.integers(-2, 3, gy.shape)                   # ±1 px in 4-px cells: every quad stays convex
    jx[:, 0] = jx[:, -1] = 0; jy[0, :] = jy[-1, :] = 0
    x = (gx + jx).astype(f32) / f32(w) - f32(1); y = (gy + jy).astype(f32) / f32(h) - f32(1)
    z = g.choice(np.array([0.0, 0.25, 0.5, 1.0], f32), gx.shape) if persp else np.zeros(gx.shape, f32)
    col = g.integers(0, 5, gx.shape + (3,)).astype(f32) / f32(4)
    if persp:                                             # mvp below makes w_clip = z + 1: pre-multiply so x/w, y/w stay on the lattice
        x = x * (z + f32(1)); y = y * (z + f32(1))
    verts = np.concatenate([x[..., None], y[..., None], z[..., None], col], 2).reshape(-1, 6).astype(f32)
    i = (np.arange(ny)[:, None] * (nx + 1) + np.arange(nx)[None, :]).reshape(-1)
    flip = g.integers(0, 2, i.shape).astype(bool)         # either diagonal per cell
    a, b, c, d = i, i + 1, i + nx + 1, i + nx + 2
    t1 = np.where(flip[:, None], np.stack([a, b, d], 1), np.stack([a, b, c], 1))
    t2 = np.where(flip[:, None], np.stack([a, d, c], 1), np.stack([b, d, c], 1))
    tris = np.concatenate([t1, t2]).astype(np.uint32)
    g.shuffle(tris)
    mvp = np.eye(4, dtype=f32)
    if persp:
        mvp[3] = [0, 0, 1, 1]
    from retrofire_b200 import mathx as mx
    ctx = rf.Context(face_cull=None, depth_test=None)
    sc = scenes.Scene(f"grid_{w}x{h}_{int(persp)}", w, h, rf.FMT_RGBA8888, True, ctx,
                      [rf.DrawCall.make(tris, verts, rf.shader.new(rf.VS_MVP, rf.FS_COLOR3F), mvp, mx.viewport((0, h), (w, 0)), ctx)])
    got = run_gpu(device, sc)
    want = run_oracle(oracle, sc)
    assert want[2].frags.i >= w * h - 50 and want[2].frags.i != w * h, want[2]   # nearly, but not exactly, once per pixel
    assert_parity(got, want, name=sc.name)


def ndc_soup(n, w, h, seed, persp, extent=1.0):
    """Random triangles given directly in NDC (identity matrix, or w = z + 1), spanning up to the whole target whatever its aspect."""
    from retrofire_b200 import mathx as mx
    g = np.random.default_rng(seed)
    c = g.uniform(-1.1, 1.1, (n, 1, 2)).astype(f32)
    xy = c + g.uniform(-extent, extent, (n, 3, 2)).astype(f32) * g.uniform(0.02, 1.0, (n, 1, 1)).astype(f32)
    z = g.uniform(-0.6 if persp else -1.1, 1.1, (n, 3, 1)).astype(f32)
    if persp:
        xy = xy * (z + f32(1))
    attr = g.uniform(0, 1, (n, 3, 3)).astype(f32)
    verts = np.concatenate([xy, z, attr], 2).reshape(3 * n, -1).astype(f32)
    tris = np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
    mvp = np.eye(4, dtype=f32)
    if persp:
        mvp[3] = [0, 0, 1, 1]
    ctx = rf.Context(face_cull=None)
    return scenes.Scene(f"ndc_soup_{w}x{h}_{seed}", w, h, rf.FMT_RGBA8888, True, ctx,
                        [rf.DrawCall.make(tris, verts, rf.shader.new(rf.VS_MVP, rf.FS_COLOR3F), mvp, mx.viewport((0, h), (w, 0)), ctx)])


@pytest.mark.parametrize("persp", [False, True])
@pytest.mark.parametrize("size", [(3000, 40), (40, 3000), (4099, 33), (1, 2500)])
def test_extreme_aspect_targets(device, oracle, size, persp):
    """Very wide targets (spans crossing up to 128 tile columns: one checkpoint per column start) and very tall ones (triangles cut
    into up to 94 chunks of 32 rows), with triangles up to the size of the target."""
    w, h = size
    sc = ndc_soup(300, w, h, seed=w + h + int(persp), persp=persp)
    want = run_oracle(oracle, sc)
    assert want[2].frags.i > 4 * w * h, want[2]
    assert_parity(run_gpu(device, sc), want, name=sc.name)


def test_nan_in_shader_max_is_ignored(device, oracle):
    """`f32::max` drops a NaN argument (crates.rs:44, solids.rs:75); the oracle side is pinned in tests/test_oracle_golden.py."""
    from tests.test_oracle_golden import nan_max_scenes
    for sc in nan_max_scenes():
        check(device, oracle, sc)


def test_nan_and_inf_vertices_behave_like_the_reference(device, oracle):
    """NaN / infinite positions and attributes: outcodes treat NaN as inside (`d > 0.0` is false), saturating
    casts map NaN to 0 rows — the CUDA path must make exactly the oracle's decisions."""
    sc = scenes.random_soup(2000, 320, 240, seed=61, lanes_kind="color3", big=True)
    v = sc.draws[0].verts
    g = np.random.default_rng(3)
    for k, val in enumerate((np.nan, np.inf, -np.inf)):
        rows = g.choice(v.shape[0], 40, replace=False)
        cols = g.integers(0, v.shape[1], 40)
        v[rows, cols] = val
    try:
        want = run_oracle(oracle, sc)
    except rf.RetrofireError as e:      # the reference would panic (span outside the target): the ABI must report the same
        with pytest.raises(rf.RetrofireError) as ge:
            run_gpu(device, sc)
        assert ge.value.status == e.status
        return
    assert_parity(run_gpu(device, sc), want, name="nan-inf")


@pytest.mark.parametrize("persp", [False, True])
def test_object_culling_ties(device, oracle, persp):
    """`BBox::visibility` (scene.rs:59-87) decided on ties: 400 objects whose boxes have corners on a lattice, many touching a
    clip plane exactly from outside or inside, degenerate (flat or point) boxes, boxes behind w = 0. Frame, Stats and objs.o must
    equal the oracle's; which objects are skipped is visible in `calls` and `prims.i`."""
    import dataclasses
    from retrofire_b200 import mathx as mx
    g = np.random.default_rng(91 + int(persp))
    w, h = 96, 64
    mvp = np.eye(4, dtype=f32)
    if persp:
        mvp[3] = [0, 0, 1, 1]
    ctx = rf.Context(face_cull=None)
    shd = rf.shader.new(rf.VS_MVP, rf.FS_COLOR3F)
    vp = mx.viewport((0, h), (w, 0))
    draws = []
    for k in range(400):
        lo = g.integers(-6, 6, 3).astype(f32) / f32(4)                       # -1.5 .. 1.25 in quarters
        ext = g.integers(0, 4, 3).astype(f32) / f32(4) * g.integers(0, 2, 3).astype(f32)   # some axes flat
        hi = lo + ext
        # geometry strictly inside its box (the reference trusts the box), two triangles
        t = g.uniform(0, 1, (6, 3)).astype(f32)
        pos = lo + (hi - lo) * t
        verts = np.concatenate([pos, g.uniform(0, 1, (6, 3)).astype(f32)], 1).astype(f32)
        d = rf.DrawCall.make(np.array([[0, 1, 2], [3, 4, 5]], np.uint32), verts, shd, mvp, vp, ctx)
        draws.append(dataclasses.replace(d, bbox=np.stack([lo, hi]).astype(f32)))
    sc = scenes.Scene(f"bbox_ties_{int(persp)}", w, h, rf.FMT_RGBA8888, True, ctx, draws)
    got, want = run_gpu(device, sc), run_oracle(oracle, sc)
    assert 20 < want[2].objs.o < 380, want[2]
    # independent of the oracle's C++: numpy outcodes of the eight corners (exact on this lattice)
    assert want[2].objs.o == sum(not scenes._bbox_hidden(d.bbox, mvp) for d in draws)
    assert (got[2].objs.i, got[2].objs.o) == (want[2].objs.i, want[2].objs.o) == (400, want[2].objs.o)
    assert_parity(got, want, name=sc.name)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_api_sequence_fuzz_persistent_targets(device, oracle, seed):
    """What a frame loop does to the queue (front/src/minifb.rs:114-152, crates.rs:98-133): three persistent targets of different
    sizes and formats, and a random sequence of partial clears (colour only, depth only, both), draws into any of them in any
    interleaving (a clear may follow draws of the same pass), flushes, stats reads and downloads. Each target is mirrored by an
    oracle target that receives the same operations in submission order; every download must match, and so must the counters."""
    g = np.random.default_rng(seed)
    specs = [(96, 64, rf.FMT_RGBA8888, True), (70, 130, rf.FMT_XRGB8888, True), (33, 40, rf.FMT_RGB565, False)]
    fbs = [device.framebuf(w, h, fmt, dep) for (w, h, fmt, dep) in specs]
    refs = [oracle.HostTarget(w, h, fmt, dep) for (w, h, fmt, dep) in specs]
    kinds = ["color3", "uv", "disc", "lit", "checker"]
    want_stats = rf.Stats()
    try:
        device.stats(reset=True)
        for step in range(120):
            i = int(g.integers(0, 3)); w, h, fmt, dep = specs[i]
            op = g.integers(0, 10)
            if op < 2:
                ctx = rf.Context(color_clear=None if g.integers(0, 3) == 0 else tuple(int(v) for v in g.integers(0, 256, 4)),
                                 depth_clear=None if g.integers(0, 3) == 0 else float(g.choice([np.inf, 4.0, 1.0])))
                fbs[i].clear(ctx)
                refs[i].clear(ctx.color_clear, ctx.depth_clear)
            elif op < 8:
                ctx = rf.Context(face_cull=[None, rf.FaceCull.Back][g.integers(0, 2)], depth_test=[None, rf.Ordering.Less][g.integers(0, 2)] if dep else None,
                                 depth_write=bool(g.integers(0, 3)), color_write=bool(g.integers(0, 5)))
                sc = scenes.random_soup(int(g.integers(1, 60)), w, h, seed=int(g.integers(1, 1 << 30)), lanes_kind=kinds[g.integers(0, len(kinds))],
                                        big=bool(g.integers(0, 2)), ctx=ctx)
                for d in sc.draws:
                    device.render(d, fbs[i])
                    want_stats += oracle.render(d, refs[i])
            elif op == 8:
                device.flush() if hasattr(device, "flush") else device.sync()
            else:
                assert np.array_equal(fbs[i].download_color(), refs[i].host_color()), (step, i)
                if dep:
                    assert depth_equal(fbs[i].download_depth(), refs[i].depth), (step, i)
        for i, (fb, ref) in enumerate(zip(fbs, refs)):
            assert np.array_equal(fb.download_color(), ref.host_color()), i
            if specs[i][3]:
                assert depth_equal(fb.download_depth(), ref.depth), i
        assert device.stats(reset=True).counters() == want_stats.counters()
    finally:
        for fb in fbs:
            fb._destroy(); device._targets.remove(fb)


def test_arena_growth_replays_the_pass(device, oracle):
    """A pass that overflows the span/piece arenas is replayed transparently after growth."""
    sc = scenes.random_soup(6000, 1920, 1080, seed=31, lanes_kind="lit", big=True)
    check(device, oracle, sc)


@pytest.mark.parametrize("kind", ["soup", "lines"])
def test_row_bands_of_any_height_tile_the_frame(device, oracle, kind):
    """Bands that ignore the 32-row tile grid — one row, a tile boundary minus one, empty tail — each equal to the oracle's band,
    and together equal to the unsharded frame (pixels and fragment counters)."""
    w, h = 200, 150
    sc = scenes.random_soup(1500, w, h, seed=77, lanes_kind="lit", big=True) if kind == "soup" else scenes.random_lines(1500, w, h, seed=78)
    whole = run_oracle(oracle, sc)
    cuts = [0, 1, 31, 32, 33, 97, 98, 149, 150]
    color = np.zeros_like(whole[0]); depth = np.zeros_like(whole[1]); fi = fo = 0
    for y0, y1 in zip(cuts[:-1], cuts[1:]):
        device.set_row_band(y0, y1)
        try:
            got = run_gpu(device, sc)
        finally:
            device.set_row_band(0, 0xFFFFFFFF)
        want = run_oracle(oracle, sc, band=(y0, y1))
        band = lambda r: (r[0][y0:y1], r[1][y0:y1], r[2])
        assert_parity(band(got), band(want), name=f"band {y0}:{y1}")
        assert not got[0][:y0].any() and not got[0][y1:].any(), "rows outside the band must stay untouched"
        color[y0:y1] = got[0][y0:y1]; depth[y0:y1] = got[1][y0:y1]
        fi += got[2].frags.i; fo += got[2].frags.o
    assert (color == whole[0]).all() and depth_equal(depth, whole[1])
    assert (fi, fo) == (whole[2].frags.i, whole[2].frags.o)


def test_fuzz_random_frames_through_one_context(device, oracle):
    """Seeded fuzz: 100 frames of random size (odd sizes included), pixel format, Context flags and draw mix (triangle soups of
    every lane layout, lines, an indexed mesh, sprites, depth-sorted draws), all through the same context so that arenas,
    slots and work lists are reused in every state the previous frame left them in."""
    import dataclasses
    g = np.random.default_rng(int(os.environ.get("RF_FUZZ_SEED", "2026")))  # other seeds: scratch/emu_fuzz.sh
    kinds = ["color3", "uv", "disc", "lit", "color4", "checker", "normal", "texclamp", "lanes8"]
    fmts = [rf.FMT_RGBA8888, rf.FMT_XRGB8888, rf.FMT_ARGB8888, rf.FMT_BGRA8888, rf.FMT_RGB888, rf.FMT_RGB565, rf.FMT_RGBA4444]
    for frame in range(int(os.environ.get("RF_FUZZ_FRAMES", "100"))):
        w, h = int(g.integers(33, 700)), int(g.integers(33, 420))
        ctx = rf.Context(face_cull=[None, rf.FaceCull.Back, rf.FaceCull.Front][g.integers(0, 3)],
                         depth_test=[None, rf.Ordering.Less, rf.Ordering.Less, rf.Ordering.Greater][g.integers(0, 4)],
                         depth_write=bool(g.integers(0, 4) > 0), color_write=bool(g.integers(0, 8) > 0),
                         depth_sort=[None, None, rf.DepthSort.BackToFront, rf.DepthSort.FrontToBack][g.integers(0, 4)])
        if ctx.depth_test == rf.Ordering.Greater:
            ctx.depth_clear = 0.001
        draws = []
        for _ in range(int(g.integers(1, 5))):
            what = g.integers(0, 10)
            seed = int(g.integers(1, 1 << 30))
            if what < 6:
                n = int(g.integers(1, 2500))
                draws += scenes.random_soup(n, w, h, seed=seed, lanes_kind=kinds[g.integers(0, len(kinds))], big=bool(g.integers(0, 2)), ctx=ctx).draws
            elif what < 8:
                draws += scenes.random_lines(int(g.integers(1, 800)), w, h, seed=seed, ctx=ctx).draws
            elif what == 8:
                d = scenes.bunny(subdiv=0, w=w, h=h).draws[0]
                draws.append(dataclasses.replace(d, face_cull=ctx.face_cull or 0, depth_test=ctx.depth_test or 0, color_write=ctx.color_write,
                                                 depth_write=ctx.depth_write, depth_sort=0 if ctx.depth_sort is None else int(ctx.depth_sort)))
            else:
                d = scenes.sprites(int(g.integers(10, 400)), w=w, h=h).draws[0]
                draws.append(dataclasses.replace(d, face_cull=ctx.face_cull or 0, depth_test=ctx.depth_test or 0, color_write=ctx.color_write,
                                                 depth_write=ctx.depth_write))
        sc = scenes.Scene(f"fuzz-{frame}", w, h, fmts[g.integers(0, len(fmts))], bool(g.integers(0, 5) > 0), ctx, draws)
        try:
            want = run_oracle(oracle, sc)
        except rf.RetrofireError as e:
            # a frame the reference would panic on (a span outside the target, render/target.rs:148,173): the same status must
            # come back from the device path, and the context must keep working for the frames after it
            with pytest.raises(rf.RetrofireError) as ge:
                run_gpu(device, sc)
            assert ge.value.status == e.status, (sc.name, ge.value, e)
            continue
        assert_parity(run_gpu(device, sc), want, name=sc.name)


def test_generated_nan_depth_has_the_host_bit_pattern(device, oracle):
    """The NaN contract. A trapezoid half of zero height makes `recip_dy = inf` and `dl = 0 * inf = NaN` (raster.rs:263-270);
    with `depth_test = None` every fragment writes its z, so the NaN lands in the depth buffer (round 1's hardware failure: the same
    pixels, but x86 generates 0xFFC00000 and CUDA 0x7FFFFFFF). The device writes the host's pattern for NaNs it generates."""
    sc = lattice_scene((64, 32), False, "none")  # round 1's failing case: six NaN depth values
    want = run_oracle(oracle, sc)
    got = run_gpu(device, sc)
    assert_parity(got, want, name=sc.name)
    nan = np.isnan(want[1])
    assert nan.any(), "the scene must put generated NaNs into the depth buffer"
    assert np.array_equal(np.isnan(got[1]), nan)
    assert (got[1].view(np.uint32)[nan] == 0xFFC00000).all(), "device-generated NaN depth must carry the x86 default-NaN bits"
    assert (want[1].view(np.uint32)[nan] == 0xFFC00000).all()
