//! Safe shim: `retrofire_core::render::render` / `Batch` with the same argument meaning, executed by
//! librf_b200.so. Swap `re::core::render::{render, Batch}` for `retrofire_b200::{render, Batch}` and
//! the frame's target for a `GpuTarget` (SURVEY §8b "Who calls it").
//!
//! NOT compiled in the build container (no cargo/rustc); written against retrofire-core 0.4.0.
//!
//! * User closures cannot cross the FFI, so the shader argument is a value implementing `GpuShader`
//!   (a pair from the fixed catalogue) instead of `shader::new(vs_fn, fs_fn)`.
//! * `Vertex` and tuples are `repr(Rust)`: vertices are flattened field by field through the `Lanes`
//!   trait, never transmuted (reference: core/src/geom/prim.rs:18-22).
//! * A non-zero `rf_status` becomes `panic!`, which is how the reference reports the same conditions
//!   (render/prim.rs:17-19, render/target.rs:148,173-174, render/tex.rs:230-231).
use retrofire_b200_sys as sys;
use retrofire_core::geom::{Edge, Mesh, Tri, Vertex, Vertex3};
use retrofire_core::math::{Color3f, Color4, Mat4, Normal3, Point3, TexCoord, Vec2};
use retrofire_core::render::ctx::DepthSort;
use retrofire_core::render::{Context, Ndc, Screen, Stats};
use std::borrow::Borrow;
use std::cell::RefCell;
use std::ffi::CStr;
use std::ptr;

/// Flattening of a varying/attribute type into f32 lanes plus the per-lane z_div mask
/// (math/vary.rs:10-15: f32/Vector/Point lanes divide, Color lanes and () do not).
pub trait Lanes {
    const N: usize;
    const PERSP_MASK: u32;
    fn write(&self, out: &mut [f32]);
}
impl Lanes for () { const N: usize = 0; const PERSP_MASK: u32 = 0; fn write(&self, _: &mut [f32]) {} }
impl Lanes for f32 { const N: usize = 1; const PERSP_MASK: u32 = 1; fn write(&self, o: &mut [f32]) { o[0] = *self; } }
impl<Sp> Lanes for Color3f<Sp> { const N: usize = 3; const PERSP_MASK: u32 = 0; fn write(&self, o: &mut [f32]) { o[..3].copy_from_slice(&self.0); } }
impl Lanes for Normal3 { const N: usize = 3; const PERSP_MASK: u32 = 0b111; fn write(&self, o: &mut [f32]) { o[..3].copy_from_slice(&self.0); } }
impl Lanes for TexCoord { const N: usize = 2; const PERSP_MASK: u32 = 0b11; fn write(&self, o: &mut [f32]) { o[..2].copy_from_slice(&self.0); } }
impl<B> Lanes for Vec2<B> { const N: usize = 2; const PERSP_MASK: u32 = 0b11; fn write(&self, o: &mut [f32]) { o[..2].copy_from_slice(&self.0); } }
impl<T: Lanes, U: Lanes> Lanes for (T, U) {
    const N: usize = T::N + U::N;
    const PERSP_MASK: u32 = T::PERSP_MASK | (U::PERSP_MASK << T::N);
    fn write(&self, o: &mut [f32]) { self.0.write(&mut o[..T::N]); self.1.write(&mut o[T::N..]); }
}

/// A (vertex, fragment) pair of the catalogue plus its uniforms; `Uni` is what the reference passes as `uniform`.
pub trait GpuShader<Uni> {
    const VS: u32;
    const FS: u32;
    fn vs_uniform(uni: &Uni) -> [f32; sys::RF_VS_UNIFORM_F32];
    fn fs_uniform(&self) -> [f32; sys::RF_FS_UNIFORM_F32] { [0.0; sys::RF_FS_UNIFORM_F32] }
    fn texture(&self) -> *const sys::rf_texture { ptr::null() }
}

type Proj<B> = retrofire_core::math::ProjMat3<B>;
type M4<S, D> = Mat4<S, D>;

// ---- the catalogue: one value type per (vertex, fragment) shader pair of the demos/tests ----------------
// The numeric ids are `rf_vs_id` / `rf_fs_id` of include/retrofire_b200.h.

macro_rules! mvp_shader {
    ($(#[$doc:meta])* $name:ident, $vs:expr, $fs:expr) => {
        $(#[$doc])*
        pub struct $name;
        impl<B> GpuShader<&Proj<B>> for $name {
            const VS: u32 = $vs;
            const FS: u32 = $fs;
            fn vs_uniform(m: &&Proj<B>) -> [f32; 32] { mat_uniform(&m.0, None) }
        }
    };
}
mvp_shader!(
    /// `|v, mvp| vertex(mvp.apply(&v.pos), v.attrib)` + `|f| f.var.to_color4()` (hello_tri.rs:22-26 non-fp, debug.rs:21-31).
    MvpColor3f, 0, 0);
mvp_shader!(
    /// hello_tri.rs:13-18 with the `fp` feature: attrib -> `powf(c, 2.2)` per vertex, `powf(c, 1/2.2)` per fragment (±1 LSB).
    MvpLinearizeColor3fSrgb, 1, 1);
mvp_shader!(
    /// Four colour lanes passed through (render/debug.rs:34-38, demos/wasm/src/triangle.rs:41).
    MvpColor4f, 0, 2);
mvp_shader!(
    /// Checker floor of crates.rs:32-36: `(uv.x() > 0.5) ^ (uv.y() > 0.5) ? 0.8 : 0.1` gray.
    MvpChecker, 0, 3);
mvp_shader!(
    /// curses.rs:49-56: normal visualisation `n / 2 + 0.5`.
    MvpNormalVis, 0, 8);

/// solids.rs:70-83: uniform `(&mvp, &spin)`; per-vertex diffuse term from the spun normal, colour from the model normal.
pub struct SolidsColor3f;
impl<B, S, D> GpuShader<(&Proj<B>, &M4<S, D>)> for SolidsColor3f {
    const VS: u32 = 2; // RF_VS_SOLIDS
    const FS: u32 = 0; // RF_FS_COLOR3F
    fn vs_uniform(u: &(&Proj<B>, &M4<S, D>)) -> [f32; 32] { mat_uniform(&u.0.0, Some(&u.1.0)) }
}

/// sprites.rs:40-52: uniform `(&modelview, &project)`; view-space billboard offset 0.008 * corner; disc with `discard`.
pub struct SpriteDisc;
impl<S, D, B> GpuShader<(&M4<S, D>, &Proj<B>)> for SpriteDisc {
    const VS: u32 = 3; // RF_VS_SPRITE
    const FS: u32 = 7; // RF_FS_SPRITE_DISC
    fn vs_uniform(u: &(&M4<S, D>, &Proj<B>)) -> [f32; 32] { mat_uniform(&u.0.0, Some(&u.1.0)) }
}

/// tests/rendering.rs:27-30, square.rs:39-40, hello.rs:33-36: `SamplerClamp.sample(&tex, uv)`.
pub struct MvpTexClamp<'t> { pub tex: &'t GpuTexture<'t> }
impl<B> GpuShader<&Proj<B>> for MvpTexClamp<'_> {
    const VS: u32 = 0;
    const FS: u32 = 5; // RF_FS_TEX_CLAMP
    fn vs_uniform(m: &&Proj<B>) -> [f32; 32] { mat_uniform(&m.0, None) }
    fn texture(&self) -> *const sys::rf_texture { self.tex.t }
}

/// benches/fill.rs:74-91: `SamplerRepeatPot` (power-of-two sizes only; anything else is RF_E_BAD_TEXTURE = the reference's assert).
pub struct MvpTexRepeatPot<'t> { pub tex: &'t GpuTexture<'t> }
impl<B> GpuShader<&Proj<B>> for MvpTexRepeatPot<'_> {
    const VS: u32 = 0;
    const FS: u32 = 6; // RF_FS_TEX_REPEAT_POT
    fn vs_uniform(m: &&Proj<B>) -> [f32; 32] { mat_uniform(&m.0, None) }
    fn texture(&self) -> *const sys::rf_texture { self.tex.t }
}

/// `SamplerOnce.sample(&tex, uv)` (render/tex.rs:313-357): coordinates are assumed inside the texture; outside it the
/// reference panics (slice index) and this path reports RF_E_BAD_TEXTURE, which `check` turns into the same panic.
pub struct MvpTexOnce<'t> { pub tex: &'t GpuTexture<'t> }
impl<B> GpuShader<&Proj<B>> for MvpTexOnce<'_> {
    const VS: u32 = sys::RF_VS_MVP;
    const FS: u32 = sys::RF_FS_TEX_ONCE;
    fn vs_uniform(m: &&Proj<B>) -> [f32; 32] { mat_uniform(&m.0, None) }
    fn texture(&self) -> *const sys::rf_texture { self.tex.t }
}

/// crates.rs:39-47: varying `(Normal3, TexCoord)`, `kd = lerp(max(n·l, 0), 0.4, 1.0)`, texel * kd.
pub struct MvpTexClampLit<'t> { pub tex: &'t GpuTexture<'t>, pub light_dir: Normal3 }
impl<B> GpuShader<&Proj<B>> for MvpTexClampLit<'_> {
    const VS: u32 = 0;
    const FS: u32 = 4; // RF_FS_TEX_CLAMP_LIT
    fn vs_uniform(m: &&Proj<B>) -> [f32; 32] { mat_uniform(&m.0, None) }
    fn fs_uniform(&self) -> [f32; sys::RF_FS_UNIFORM_F32] {
        let mut u = [0.0; sys::RF_FS_UNIFORM_F32];
        u[..3].copy_from_slice(&self.light_dir.0);
        u
    }
    fn texture(&self) -> *const sys::rf_texture { self.tex.t }
}

fn mat_uniform(a: &[[f32; 4]; 4], b: Option<&[[f32; 4]; 4]>) -> [f32; 32] {
    let mut u = [0.0f32; 32];
    for r in 0..4 { u[4 * r..4 * r + 4].copy_from_slice(&a[r]); }
    if let Some(b) = b { for r in 0..4 { u[16 + 4 * r..20 + 4 * r].copy_from_slice(&b[r]); } }
    u
}

/// One GPU (rf_ctx). Single-owner like `Context` (`RefCell` inside makes it !Sync).
pub struct Gpu { ctx: *mut sys::rf_ctx, _not_sync: RefCell<()> }
impl Gpu {
    pub fn new(device: i32) -> Self {
        let mut ctx = ptr::null_mut();
        let st = unsafe { sys::rf_ctx_create(device, ptr::null_mut(), &mut ctx) };
        assert!(st == sys::RF_OK, "rf_ctx_create failed ({st}): no sm_100 GPU?");
        Gpu { ctx, _not_sync: RefCell::new(()) }
    }
    fn check(&self, st: sys::rf_status) {
        if st != sys::RF_OK {
            let msg = unsafe { CStr::from_ptr(sys::rf_last_error(self.ctx)) }.to_string_lossy().into_owned();
            panic!("retrofire-b200: status {st}: {msg}"); // the reference panics in the same situations
        }
    }
}
impl Drop for Gpu { fn drop(&mut self) { unsafe { sys::rf_ctx_destroy(self.ctx) } } }

/// Device-resident `Framebuf<Colorbuf<_, Fmt>, Buf2<f32>>` (render/target.rs:35-58).
pub struct GpuTarget<'g> { gpu: &'g Gpu, t: *mut sys::rf_target, pub w: u32, pub h: u32 }
impl<'g> GpuTarget<'g> {
    pub fn new(gpu: &'g Gpu, w: u32, h: u32, fmt: u32, depth: bool) -> Self {
        let mut t = ptr::null_mut();
        gpu.check(unsafe { sys::rf_target_create(gpu.ctx, w, h, fmt, depth as i32, &mut t) });
        GpuTarget { gpu, t, w, h }
    }
    /// `Frame::clear` (front/src/lib.rs:103-120).
    pub fn clear(&mut self, ctx: &Context) {
        let rgba = ctx.color_clear.map(|c: Color4| c.0);
        let z = ctx.depth_clear.map(|d| d.recip());
        self.gpu.check(unsafe {
            sys::rf_target_clear(self.gpu.ctx, self.t, rgba.as_ref().map_or(ptr::null(), |c| c.as_ptr()), z.as_ref().map_or(ptr::null(), |z| z as *const f32))
        });
    }
    /// Draws into `self.queued()` are only queued (no device round trip per `render()`: a crates frame is 1,090 calls). They
    /// execute at the next `finish` / `download_into`; their `Stats` are added to `ctx.stats` by `finish`.
    pub fn queued(&mut self) -> Queued<'_, 'g> { Queued(self) }
    /// Executes everything queued and adds the `Stats` of those draws to `ctx.stats` (render.rs:206 does this per call).
    pub fn finish(&mut self, ctx: &Context) {
        let mut st = sys::rf_stats::default();
        self.gpu.check(unsafe { sys::rf_ctx_stats(self.gpu.ctx, &mut st, 1) });
        *ctx.stats.borrow_mut() += stats_from(&st);
    }
    /// Present: copy the colour buffer into a host `Buf2`-like slice honouring its stride (util/buf.rs:437-439).
    pub fn download_into<T>(&mut self, data: &mut [T], stride: usize) {
        assert!(data.len() >= stride * (self.h as usize - 1) + self.w as usize);
        self.gpu.check(unsafe { sys::rf_target_download_color(self.gpu.ctx, self.t, data.as_mut_ptr().cast(), stride) });
    }
}
impl Drop for GpuTarget<'_> { fn drop(&mut self) { unsafe { sys::rf_target_destroy(self.t) } } }

/// Device copy of a `Texture<Buf2<Color3>>` / `Texture<Buf2<Color4>>` (render/tex.rs:33-37,190-213). Nearest sampling only,
/// as in the reference; `w`/`h` are the float dimensions the samplers multiply by.
pub struct GpuTexture<'g> { _gpu: &'g Gpu, t: *mut sys::rf_texture }
impl<'g> GpuTexture<'g> {
    /// `texels`: row-major, `stride` elements per row (util/buf.rs:437-439); `texel_fmt` = `sys::RF_TEXEL_RGB888` (`Color3`
    /// texels, 3 bytes) or `sys::RF_TEXEL_RGBA8888` (`Color4`, 4 bytes) — rf_texel_fmt, NOT the rf_pixel_fmt of targets.
    pub fn new<T>(gpu: &'g Gpu, w: u32, h: u32, texel_fmt: u32, texels: &[T], stride: usize) -> Self {
        assert!(texel_fmt == sys::RF_TEXEL_RGB888 || texel_fmt == sys::RF_TEXEL_RGBA8888, "texel_fmt is an rf_texel_fmt");
        let fmt = texel_fmt;
        assert!(texels.len() >= stride * (h as usize).saturating_sub(1) + w as usize);
        let mut t = ptr::null_mut();
        gpu.check(unsafe { sys::rf_texture_create(gpu.ctx, w, h, fmt, texels.as_ptr().cast(), stride, &mut t) });
        GpuTexture { _gpu: gpu, t }
    }
}
impl Drop for GpuTexture<'_> { fn drop(&mut self) { unsafe { sys::rf_texture_destroy(self.t) } } }

/// What `render()` draws into. `GpuTarget` itself behaves like the reference (the draw has executed and `ctx.stats` is updated
/// when `render()` returns: one flush + wait per call); `Queued` only records the draw.
pub trait GpuTgt {
    const SYNC: bool;
    fn gpu(&self) -> &Gpu;
    fn raw(&mut self) -> *mut sys::rf_target;
}
impl GpuTgt for GpuTarget<'_> {
    const SYNC: bool = true;
    fn gpu(&self) -> &Gpu { self.gpu }
    fn raw(&mut self) -> *mut sys::rf_target { self.t }
}
/// `target.queued()`: see `GpuTarget::queued`.
pub struct Queued<'a, 'g>(pub &'a mut GpuTarget<'g>);
impl GpuTgt for Queued<'_, '_> {
    const SYNC: bool = false;
    fn gpu(&self) -> &Gpu { self.0.gpu }
    fn raw(&mut self) -> *mut sys::rf_target { self.0.t }
}
impl<T: GpuTgt> GpuTgt for &mut T {
    const SYNC: bool = T::SYNC;
    fn gpu(&self) -> &Gpu { (**self).gpu() }
    fn raw(&mut self) -> *mut sys::rf_target { (**self).raw() }
}

/// Persistent device copy of a primitive list and its vertices (rf_mesh): what `Batch` clones per call (batch.rs:62-84)
/// is uploaded once. Draw it with `render_mesh`.
pub struct GpuMesh<'g, Prim, A> { gpu: &'g Gpu, m: *mut sys::rf_mesh, _p: std::marker::PhantomData<(Prim, A)> }
impl<'g, Prim: GpuPrim, A: Lanes> GpuMesh<'g, Prim, A> {
    pub fn new<Vtx: GpuVertex<Attr = A>>(gpu: &'g Gpu, prims: impl AsRef<[Prim]>, verts: impl AsRef<[Vtx]>) -> Self {
        let (flat, idx, stride) = flatten(prims.as_ref(), verts.as_ref());
        let mut m = ptr::null_mut();
        gpu.check(unsafe {
            sys::rf_mesh_create(gpu.ctx, flat.as_ptr(), verts.as_ref().len() as u32, stride as u32, idx.as_ptr(), prims.as_ref().len() as u32, Prim::KIND as u32, &mut m)
        });
        GpuMesh { gpu, m, _p: std::marker::PhantomData }
    }
}
impl<Prim, A> Drop for GpuMesh<'_, Prim, A> { fn drop(&mut self) { let _ = self.gpu; unsafe { sys::rf_mesh_destroy(self.m) } } }

/// The primitive kinds `render()` accepts (`Render for Tri<usize>` prim.rs:17-39, `Render for Edge<usize>` prim.rs:41-60).
pub trait GpuPrim: Clone {
    const KIND: u8;
    const ARITY: usize;
    fn indices(&self, out: &mut Vec<u32>);
}
fn idx32(i: usize) -> u32 { u32::try_from(i).expect("vertex index >= 2^32") }
impl GpuPrim for Tri<usize> {
    const KIND: u8 = sys::RF_PRIM_TRIS;
    const ARITY: usize = 3;
    fn indices(&self, out: &mut Vec<u32>) { out.extend(self.0.iter().map(|&i| idx32(i))); }
}
impl GpuPrim for Edge<usize> {
    const KIND: u8 = sys::RF_PRIM_EDGES;
    const ARITY: usize = 2;
    fn indices(&self, out: &mut Vec<u32>) { out.push(idx32(self.0)); out.push(idx32(self.1)); }
}

/// The vertex types `render()` accepts: `Vertex<Point3<B>, A>` (= `Vertex3<A, B>`) with an attribute that flattens into lanes.
/// `Vertex` is `repr(Rust)` (geom/prim.rs:18-22): it is read field by field, never transmuted.
pub trait GpuVertex: Clone {
    type Attr: Lanes;
    fn pos(&self) -> [f32; 3];
    fn attrib(&self) -> &Self::Attr;
}
impl<B, A: Lanes + Clone> GpuVertex for Vertex<Point3<B>, A> where Point3<B>: Clone {
    type Attr = A;
    fn pos(&self) -> [f32; 3] { self.pos.0 }
    fn attrib(&self) -> &A { &self.attrib }
}

fn flatten<Prim: GpuPrim, Vtx: GpuVertex>(prims: &[Prim], verts: &[Vtx]) -> (Vec<f32>, Vec<u32>, usize) {
    let stride = 3 + <Vtx::Attr as Lanes>::N;
    let mut flat = vec![0.0f32; verts.len() * stride];
    for (v, o) in verts.iter().zip(flat.chunks_exact_mut(stride)) {
        o[..3].copy_from_slice(&v.pos());
        v.attrib().write(&mut o[3..]);
    }
    let mut idx: Vec<u32> = Vec::with_capacity(prims.len() * Prim::ARITY);
    for p in prims { p.indices(&mut idx); }
    (flat, idx, stride)
}

fn stats_from(st: &sys::rf_stats) -> Stats {
    let mut s = Stats::new();
    s.calls = st.calls as f32;
    s.objs.i = st.objs_i as usize; s.objs.o = st.objs_o as usize;
    s.prims.i = st.prims_i as usize; s.prims.o = st.prims_o as usize;
    s.verts.i = st.verts_i as usize; s.verts.o = st.verts_o as usize;
    s.frags.i = st.frags_i as usize; s.frags.o = st.frags_o as usize;
    s.time = std::time::Duration::from_nanos(st.time_ns);
    s
}

/// `render::Batch` (render/batch.rs:31-147): the same six type parameters, public fields and typestate setters — every
/// setter returns a batch with one parameter replaced (batch.rs:42-47 `update!`), so the reference's call sites compile
/// unchanged: `Batch::new().mesh(&m).shader(s).viewport(vp).context(&ctx)` once, then
/// `batch.clone().uniform(&mvp).target(&mut target).render()` per object (demos/src/bin/crates.rs:103-130).
#[derive(Clone, Debug, Default)]
pub struct Batch<Prim, Vtx, Uni, Shd, Tgt, Ctx> {
    pub prims: Vec<Prim>,
    pub verts: Vec<Vtx>,
    pub uniform: Uni,
    pub shader: Shd,
    pub viewport: Mat4<Ndc, Screen>,
    pub target: Tgt,
    pub ctx: Ctx,
}

impl Batch<(), (), (), (), (), Context> {
    /// batch.rs:50-54
    pub fn new() -> Self { Self::default() }
}

impl<Prim, Vtx, Uni, Shd, Tgt, Ctx> Batch<Prim, Vtx, Uni, Shd, Tgt, Ctx> {
    /// batch.rs:57-64: the primitives are copied into the batch.
    pub fn primitives<P: Clone>(self, prims: impl AsRef<[P]>) -> Batch<P, Vtx, Uni, Shd, Tgt, Ctx> {
        let Batch { verts, uniform, shader, viewport, target, ctx, .. } = self;
        Batch { prims: prims.as_ref().to_vec(), verts, uniform, shader, viewport, target, ctx }
    }
    /// batch.rs:69-76: the vertices are cloned into the batch.
    pub fn vertices<V: Clone>(self, verts: impl AsRef<[V]>) -> Batch<Prim, V, Uni, Shd, Tgt, Ctx> {
        let Batch { prims, uniform, shader, viewport, target, ctx, .. } = self;
        Batch { prims, verts: verts.as_ref().to_vec(), uniform, shader, viewport, target, ctx }
    }
    /// batch.rs:79-87: clones faces and vertices from a mesh.
    pub fn mesh<A: Clone>(self, mesh: &Mesh<A>) -> Batch<Tri<usize>, Vertex3<A>, Uni, Shd, Tgt, Ctx> {
        let Batch { uniform, shader, viewport, target, ctx, .. } = self;
        Batch { prims: mesh.faces.clone(), verts: mesh.verts.clone(), uniform, shader, viewport, target, ctx }
    }
    /// batch.rs:89-94
    pub fn uniform<U: Copy>(self, uniform: U) -> Batch<Prim, Vtx, U, Shd, Tgt, Ctx> {
        let Batch { prims, verts, shader, viewport, target, ctx, .. } = self;
        Batch { prims, verts, uniform, shader, viewport, target, ctx }
    }
    /// batch.rs:97-102. The shader is a value of the catalogue (`GpuShader`) instead of a pair of closures.
    pub fn shader<S>(self, shader: S) -> Batch<Prim, Vtx, Uni, S, Tgt, Ctx> {
        let Batch { prims, verts, uniform, viewport, target, ctx, .. } = self;
        Batch { prims, verts, uniform, shader, viewport, target, ctx }
    }
    /// batch.rs:105-107
    pub fn viewport(self, viewport: Mat4<Ndc, Screen>) -> Self { Batch { viewport, ..self } }
    /// batch.rs:110-113: `&mut GpuTarget` (synchronous, like the reference) or `target.queued()`.
    pub fn target<T>(self, target: T) -> Batch<Prim, Vtx, Uni, Shd, T, Ctx> {
        let Batch { prims, verts, uniform, shader, viewport, ctx, .. } = self;
        Batch { prims, verts, uniform, shader, viewport, target, ctx }
    }
    /// batch.rs:116-121
    pub fn context(self, ctx: &Context) -> Batch<Prim, Vtx, Uni, Shd, Tgt, &Context> {
        let Batch { prims, verts, uniform, shader, viewport, target, .. } = self;
        Batch { prims, verts, uniform, shader, viewport, target, ctx }
    }
}

impl<Prim, Vtx, Uni, Shd, Tgt, Ctx> Batch<Prim, Vtx, Uni, Shd, Tgt, Ctx> {
    /// batch.rs:127-146
    pub fn render(&mut self)
    where
        Prim: GpuPrim,
        Vtx: GpuVertex,
        Uni: Copy,
        Shd: GpuShader<Uni>,
        Tgt: GpuTgt,
        Ctx: Borrow<Context>,
    {
        let Self { prims, verts, shader, uniform, viewport, target, ctx } = self;
        render(prims, verts, shader, *uniform, *viewport, target, (*ctx).borrow());
    }
}

impl<Vtx, Uni, Shd, Tgt, Ctx> Batch<Edge<usize>, Vtx, Uni, Shd, Tgt, Ctx> {
    /// batch.rs:149-158
    pub fn append(&mut self, other: Self) {
        let Batch { prims, verts, .. } = other;
        let n = self.verts.len();
        let prims = prims.into_iter().map(|e| Edge(e.0 + n, e.1 + n));
        self.verts.extend(verts);
        self.prims.extend(prims)
    }
}

impl<Vtx, Uni, Shd, Tgt, Ctx> Batch<Tri<usize>, Vtx, Uni, Shd, Tgt, Ctx> {
    /// batch.rs:160-169
    pub fn append(&mut self, other: Self) {
        let Batch { prims, verts, .. } = other;
        let n = self.verts.len();
        let prims = prims.into_iter().map(|tri| tri.map(|i| i + n));
        self.verts.extend(verts);
        self.prims.extend(prims);
    }
}

/// What differs from a plain `render()`: resident geometry and the scene loop's per-object culling on the device.
#[derive(Clone, Copy, Default)]
pub struct DrawOpts<'m> {
    /// `rf_draw.mesh`: draw the resident copy instead of uploading `prims` / `verts` (which may then be empty).
    pub mesh: Option<*const sys::rf_mesh>,
    /// `BBox<Model>` low / upp (scene.rs:59-87): the draw is skipped — and not counted in `Stats` beyond `objs.i` — when
    /// `BBox::visibility(model_to_projection)` is `Hidden`; the test of crates.rs:100-122, run on the device.
    pub bbox: Option<[f32; 6]>,
    pub _life: std::marker::PhantomData<&'m ()>,
}

/// `retrofire_core::render::render` (render.rs:134-207) on the GPU.
pub fn render<Prim, Vtx, Uni: Copy, Shd, Tgt>(
    prims: impl AsRef<[Prim]>,
    verts: impl AsRef<[Vtx]>,
    shader: &Shd,
    uniform: Uni,
    to_screen: Mat4<Ndc, Screen>,
    target: &mut Tgt,
    ctx: &Context,
) where
    Prim: GpuPrim,
    Vtx: GpuVertex,
    Shd: GpuShader<Uni>,
    Tgt: GpuTgt,
{
    render_opts(prims, verts, shader, uniform, to_screen, target, ctx, DrawOpts::default())
}

/// `render()` of a resident mesh (the arrays stay on the device between calls).
pub fn render_mesh<Prim, A, Uni: Copy, Shd, Tgt>(
    mesh: &GpuMesh<'_, Prim, A>, shader: &Shd, uniform: Uni, to_screen: Mat4<Ndc, Screen>, target: &mut Tgt, ctx: &Context, bbox: Option<[f32; 6]>,
) where
    Prim: GpuPrim,
    A: Lanes + Clone,
    Shd: GpuShader<Uni>,
    Tgt: GpuTgt,
{
    let none: [Vertex<Point3<()>, A>; 0] = [];
    let noprims: [Prim; 0] = [];
    render_opts(noprims, none, shader, uniform, to_screen, target, ctx, DrawOpts { mesh: Some(mesh.m), bbox, _life: std::marker::PhantomData })
}

#[allow(clippy::too_many_arguments)]
pub fn render_opts<Prim, Vtx, Uni: Copy, Shd, Tgt>(
    prims: impl AsRef<[Prim]>,
    verts: impl AsRef<[Vtx]>,
    shader: &Shd,
    uniform: Uni,
    to_screen: Mat4<Ndc, Screen>,
    target: &mut Tgt,
    ctx: &Context,
    opts: DrawOpts<'_>,
) where
    Prim: GpuPrim,
    Vtx: GpuVertex,
    Shd: GpuShader<Uni>,
    Tgt: GpuTgt,
{
    let (prims, verts) = (prims.as_ref(), verts.as_ref());
    let (flat, idx, stride) = flatten(prims, verts);
    let resident = opts.mesh.is_some();
    let mut vp = [0.0f32; 16];
    for r in 0..4 { vp[4 * r..4 * r + 4].copy_from_slice(&to_screen.0[r]); }
    let draw = sys::rf_draw {
        indices: if resident { ptr::null() } else { idx.as_ptr() }, n_prims: if resident { 0 } else { prims.len() as u32 },
        verts: if resident { ptr::null() } else { flat.as_ptr() }, n_verts: if resident { 0 } else { verts.len() as u32 },
        vert_stride_f32: stride as u32,
        mesh: opts.mesh.unwrap_or(ptr::null()), n_attr_lanes: <Vtx::Attr as Lanes>::N as u32, persp_mask: <Vtx::Attr as Lanes>::PERSP_MASK,
        vs: Shd::VS, fs: Shd::FS, vs_uniform: Shd::vs_uniform(&uniform), fs_uniform: shader.fs_uniform(),
        texture: shader.texture(), viewport: vp,
        face_cull: match ctx.face_cull { None => sys::RF_CULL_NONE, Some(retrofire_core::render::ctx::FaceCull::Back) => sys::RF_CULL_BACK, Some(_) => sys::RF_CULL_FRONT },
        depth_test: match ctx.depth_test {
            None => sys::RF_DEPTH_NONE, Some(core::cmp::Ordering::Less) => sys::RF_DEPTH_LESS,
            Some(core::cmp::Ordering::Equal) => sys::RF_DEPTH_EQUAL, Some(_) => sys::RF_DEPTH_GREATER,
        },
        color_write: ctx.color_write as u8, depth_write: ctx.depth_write as u8,
        depth_sort: match ctx.depth_sort { None => sys::RF_SORT_NONE, Some(DepthSort::FrontToBack) => sys::RF_SORT_FRONT_TO_BACK, Some(DepthSort::BackToFront) => sys::RF_SORT_BACK_TO_FRONT },
        prim_kind: Prim::KIND, bbox_cull: opts.bbox.is_some() as u8, _pad: [0; 1], bbox: opts.bbox.unwrap_or([0.0; 6]),
    };
    let raw = target.raw();
    let gpu = target.gpu();
    if Tgt::SYNC {
        // stats_out != NULL: flush + wait — the reference's "done when render() returns", Stats into ctx.stats (render.rs:206)
        let mut st = sys::rf_stats::default();
        gpu.check(unsafe { sys::rf_render(gpu.ctx, raw, &draw, &mut st) });
        *ctx.stats.borrow_mut() += stats_from(&st);
    } else {
        // queued: the geometry is copied during the call (the borrow ends here, as in the reference); `GpuTarget::finish`
        // executes the frame's draws in one pass and folds their Stats
        gpu.check(unsafe { sys::rf_render(gpu.ctx, raw, &draw, ptr::null_mut()) });
    }
}
