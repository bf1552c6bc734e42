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
use retrofire_core::geom::{Tri, Vertex};
use retrofire_core::math::{Color3f, Color4, Mat4, Normal3, Point3, TexCoord, Vec2};
use retrofire_core::render::ctx::DepthSort;
use retrofire_core::render::{Context, Ndc, Screen, Stats};
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

/// `|v, mvp| vertex(mvp.apply(&v.pos), v.attrib)` + `|f| f.var.to_color4()` (demos solids/hello_tri non-fp).
pub struct MvpColor3f;
impl<B> GpuShader<&retrofire_core::math::ProjMat3<B>> for MvpColor3f {
    const VS: u32 = 0; // RF_VS_MVP
    const FS: u32 = 0; // RF_FS_COLOR3F
    fn vs_uniform(m: &&retrofire_core::math::ProjMat3<B>) -> [f32; 32] { mat_uniform(&m.0, None) }
}
// ... one unit struct per catalogue entry: MvpTexClamp{tex}, MvpTexClampLit{tex, light_dir}, MvpChecker,
// SolidsColor3f (uniform (&mvp, &spin)), SpriteDisc (uniform (&modelview, &proj)), MvpNormalVis, ...

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
    /// Present: copy the colour buffer into a host `Buf2`-like slice honouring its stride (util/buf.rs:437-439).
    pub fn download_into<T>(&mut self, data: &mut [T], stride: usize) {
        assert!(data.len() >= stride * (self.h as usize - 1) + self.w as usize);
        self.gpu.check(unsafe { sys::rf_target_download_color(self.gpu.ctx, self.t, data.as_mut_ptr().cast(), stride) });
    }
}
impl Drop for GpuTarget<'_> { fn drop(&mut self) { unsafe { sys::rf_target_destroy(self.t) } } }

/// `retrofire_core::render::render` (render.rs:134-207) on the GPU.
pub fn render<A, Uni: Copy, Shd>(
    prims: impl AsRef<[Tri<usize>]>,
    verts: impl AsRef<[Vertex<Point3<impl Sized>, A>]>,
    shader: &Shd,
    uniform: Uni,
    to_screen: Mat4<Ndc, Screen>,
    target: &mut GpuTarget<'_>,
    ctx: &Context,
) where
    A: Lanes,
    Shd: GpuShader<Uni>,
{
    let (prims, verts) = (prims.as_ref(), verts.as_ref());
    let stride = 3 + A::N;
    let mut flat = vec![0.0f32; verts.len() * stride];
    for (v, o) in verts.iter().zip(flat.chunks_exact_mut(stride)) {
        o[..3].copy_from_slice(&v.pos.0);
        v.attrib.write(&mut o[3..]);
    }
    let idx: Vec<u32> = prims.iter().flat_map(|t| t.0).map(|i| u32::try_from(i).expect("vertex index >= 2^32")).collect();
    let mut vp = [0.0f32; 16];
    for r in 0..4 { vp[4 * r..4 * r + 4].copy_from_slice(&to_screen.0[r]); }
    let draw = sys::rf_draw {
        indices: idx.as_ptr(), n_prims: prims.len() as u32,
        verts: flat.as_ptr(), n_verts: verts.len() as u32, vert_stride_f32: stride as u32,
        mesh: ptr::null(), n_attr_lanes: A::N as u32, persp_mask: A::PERSP_MASK,
        vs: Shd::VS, fs: Shd::FS, vs_uniform: Shd::vs_uniform(&uniform), fs_uniform: shader.fs_uniform(),
        texture: shader.texture(), viewport: vp,
        face_cull: match ctx.face_cull { None => 0, Some(retrofire_core::render::ctx::FaceCull::Back) => 1, Some(_) => 2 },
        depth_test: match ctx.depth_test { None => 0, Some(core::cmp::Ordering::Less) => 1, Some(core::cmp::Ordering::Equal) => 2, Some(_) => 3 },
        color_write: ctx.color_write as u8, depth_write: ctx.depth_write as u8,
        depth_sort: match ctx.depth_sort { None => 0, Some(DepthSort::FrontToBack) => 1, Some(DepthSort::BackToFront) => 2 }, prim_kind: 0, bbox_cull: 0, _pad: [0; 1], bbox: [0.0; 6],
    };
    let mut st = sys::rf_stats::default();
    // stats_out != NULL: flush + wait, i.e. the reference's "done when render() returns" (render.rs:206)
    target.gpu.check(unsafe { sys::rf_render(target.gpu.ctx, target.t, &draw, &mut st) });
    let mut s = Stats::new();
    s.calls = st.calls as f32;
    s.prims.i = st.prims_i as usize; s.prims.o = st.prims_o as usize;
    s.verts.i = st.verts_i as usize; s.verts.o = st.verts_o as usize;
    s.frags.i = st.frags_i as usize; s.frags.o = st.frags_o as usize;
    s.time = std::time::Duration::from_nanos(st.time_ns);
    *ctx.stats.borrow_mut() += s;
}
