//! Raw bindings, one item per declaration of `include/retrofire_b200.h` (ABI version 2).
//! NOT compiled in the build container (no cargo/rustc there); kept in lock-step with the header.
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

pub const RF_MAX_ATTR_LANES: usize = 8;
pub const RF_VS_UNIFORM_F32: usize = 32;
pub const RF_FS_UNIFORM_F32: usize = 8;
pub const RF_N_KERNELS: usize = 12;
pub const RF_ABI_VERSION: u32 = 2;

// rf_vs_id / rf_fs_id: the shader catalogue
pub const RF_VS_MVP: u32 = 0;
pub const RF_VS_MVP_LINEARIZE: u32 = 1;
pub const RF_VS_SOLIDS: u32 = 2;
pub const RF_VS_SPRITE: u32 = 3;
pub const RF_FS_COLOR3F: u32 = 0;
pub const RF_FS_COLOR3F_SRGB: u32 = 1;
pub const RF_FS_COLOR4F: u32 = 2;
pub const RF_FS_CHECKER: u32 = 3;
pub const RF_FS_TEX_CLAMP_LIT: u32 = 4;
pub const RF_FS_TEX_CLAMP: u32 = 5;
pub const RF_FS_TEX_REPEAT_POT: u32 = 6;
pub const RF_FS_SPRITE_DISC: u32 = 7;
pub const RF_FS_NORMAL_VIS: u32 = 8;
pub const RF_FS_TEX_ONCE: u32 = 9;
// rf_pixel_fmt: the colour layout of a TARGET (rf_target_create)
pub const RF_FMT_RGBA8888: u32 = 0;
pub const RF_FMT_XRGB8888: u32 = 1;
pub const RF_FMT_ARGB8888: u32 = 2;
pub const RF_FMT_BGRA8888: u32 = 3;
pub const RF_FMT_RGB888: u32 = 4;
pub const RF_FMT_RGB565: u32 = 5;
pub const RF_FMT_RGBA4444: u32 = 6;
// rf_texel_fmt: the texel layout of a TEXTURE (rf_texture_create) — a different enumeration from rf_pixel_fmt
pub const RF_TEXEL_RGB888: u32 = 0;
pub const RF_TEXEL_RGBA8888: u32 = 1;
pub const RF_PRIM_TRIS: u8 = 0;
pub const RF_PRIM_EDGES: u8 = 1;
pub const RF_SORT_NONE: u8 = 0;
pub const RF_SORT_FRONT_TO_BACK: u8 = 1;
pub const RF_SORT_BACK_TO_FRONT: u8 = 2;
pub const RF_CULL_NONE: u8 = 0;
pub const RF_CULL_BACK: u8 = 1;
pub const RF_CULL_FRONT: u8 = 2;
pub const RF_DEPTH_NONE: u8 = 0;
pub const RF_DEPTH_LESS: u8 = 1;
pub const RF_DEPTH_EQUAL: u8 = 2;
pub const RF_DEPTH_GREATER: u8 = 3;

#[repr(C)] pub struct rf_ctx { _p: [u8; 0] }
#[repr(C)] pub struct rf_target { _p: [u8; 0] }
#[repr(C)] pub struct rf_texture { _p: [u8; 0] }
#[repr(C)] pub struct rf_mesh { _p: [u8; 0] }

pub type rf_status = c_int;
pub const RF_OK: rf_status = 0;
pub const RF_E_INVALID: rf_status = 1;
pub const RF_E_INDEX_OOB: rf_status = 2;
pub const RF_E_TARGET_OOB: rf_status = 3;
pub const RF_E_BAD_TEXTURE: rf_status = 4;
pub const RF_E_UNSUPPORTED_SHADER: rf_status = 5;
pub const RF_E_CUDA: rf_status = 6;
pub const RF_E_NCCL: rf_status = 7;
pub const RF_E_NOMEM: rf_status = 8;
pub const RF_E_UNSUPPORTED: rf_status = 9;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct rf_draw {
    pub indices: *const u32,
    pub n_prims: u32,
    pub verts: *const f32,
    pub n_verts: u32,
    pub vert_stride_f32: u32,
    pub mesh: *const rf_mesh,
    pub n_attr_lanes: u32,
    pub persp_mask: u32,
    pub vs: u32,
    pub fs: u32,
    pub vs_uniform: [f32; RF_VS_UNIFORM_F32],
    pub fs_uniform: [f32; RF_FS_UNIFORM_F32],
    pub texture: *const rf_texture,
    pub viewport: [f32; 16],
    pub face_cull: u8,
    pub depth_test: u8,
    pub color_write: u8,
    pub depth_write: u8,
    pub depth_sort: u8,
    pub prim_kind: u8,
    pub bbox_cull: u8,
    pub _pad: [u8; 1],
    pub bbox: [f32; 6],
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct rf_stats {
    pub calls: u64,
    pub prims_i: u64, pub prims_o: u64,
    pub verts_i: u64, pub verts_o: u64,
    pub frags_i: u64, pub frags_o: u64,
    pub time_ns: u64,
    pub objs_i: u64, pub objs_o: u64,
}

unsafe extern "C" {
    pub fn rf_abi_version() -> u32;
    pub fn rf_ctx_create(device: c_int, stream: *mut c_void, out: *mut *mut rf_ctx) -> rf_status;
    pub fn rf_ctx_destroy(ctx: *mut rf_ctx);
    pub fn rf_last_error(ctx: *const rf_ctx) -> *const c_char;
    pub fn rf_ctx_set_row_band(ctx: *mut rf_ctx, y0: u32, y1: u32) -> rf_status;
    pub fn rf_target_create(ctx: *mut rf_ctx, w: u32, h: u32, color_fmt: u32, has_depth: c_int, out: *mut *mut rf_target) -> rf_status;
    pub fn rf_target_destroy(t: *mut rf_target);
    pub fn rf_target_clear(ctx: *mut rf_ctx, t: *mut rf_target, rgba: *const u8, depth_recip: *const f32) -> rf_status;
    pub fn rf_target_upload_color(ctx: *mut rf_ctx, t: *mut rf_target, host: *const c_void, stride_elems: usize) -> rf_status;
    pub fn rf_target_download_color(ctx: *mut rf_ctx, t: *mut rf_target, host: *mut c_void, stride_elems: usize) -> rf_status;
    pub fn rf_target_upload_depth(ctx: *mut rf_ctx, t: *mut rf_target, host: *const f32, stride_elems: usize) -> rf_status;
    pub fn rf_target_download_depth(ctx: *mut rf_ctx, t: *mut rf_target, host: *mut f32, stride_elems: usize) -> rf_status;
    pub fn rf_host_alloc(bytes: usize, out: *mut *mut c_void) -> rf_status;
    pub fn rf_host_free(p: *mut c_void);
    pub fn rf_target_download_color_async(ctx: *mut rf_ctx, t: *mut rf_target, host: *mut c_void, stride_elems: usize) -> rf_status;
    pub fn rf_target_color_devptr(t: *mut rf_target) -> *mut c_void;
    pub fn rf_target_depth_devptr(t: *mut rf_target) -> *mut c_void;
    pub fn rf_texture_create(ctx: *mut rf_ctx, w: u32, h: u32, texel_fmt: u32, data: *const c_void, stride_elems: usize, out: *mut *mut rf_texture) -> rf_status;
    pub fn rf_texture_destroy(t: *mut rf_texture);
    pub fn rf_mesh_create(ctx: *mut rf_ctx, verts: *const f32, n_verts: u32, vert_stride_f32: u32, indices: *const u32, n_prims: u32, prim_kind: u32, out: *mut *mut rf_mesh) -> rf_status;
    pub fn rf_mesh_destroy(m: *mut rf_mesh);
    pub fn rf_render(ctx: *mut rf_ctx, target: *mut rf_target, draw: *const rf_draw, stats_out: *mut rf_stats) -> rf_status;
    pub fn rf_render_frames(ctx: *mut rf_ctx, targets: *const *mut rf_target, n_frames: u32, draw: *const rf_draw, vs_uniforms: *const f32) -> rf_status;
    pub fn rf_render_many(ctx: *mut rf_ctx, target: *mut rf_target, draws: *const rf_draw, n_draws: u32) -> rf_status;
    pub fn rf_ctx_peer_export(ctx: *mut rf_ctx, ipc_handle_out: *mut u8, devptr_out: *mut *mut c_void) -> rf_status;
    pub fn rf_ctx_peer_attach(ctx: *mut rf_ctx, world: u32, rank: u32, ipc_handles: *const u8, devptrs: *const *mut c_void) -> rf_status;
    pub fn rf_target_peer_export(ctx: *mut rf_ctx, t: *mut rf_target, ipc_handle_out: *mut u8, devptr_out: *mut *mut c_void) -> rf_status;
    pub fn rf_target_peer_attach(ctx: *mut rf_ctx, t: *mut rf_target, world: u32, rank: u32, ipc_handles: *const u8, devptrs: *const *mut c_void) -> rf_status;
    pub fn rf_ctx_replays(ctx: *mut rf_ctx, out: *mut u64) -> rf_status;
    pub fn rf_ctx_set_geometry_async(ctx: *mut rf_ctx, on: c_int) -> rf_status;
    pub fn rf_flush(ctx: *mut rf_ctx) -> rf_status;
    pub fn rf_sync(ctx: *mut rf_ctx) -> rf_status;
    pub fn rf_ctx_stats(ctx: *mut rf_ctx, out: *mut rf_stats, reset: c_int) -> rf_status;
    pub fn rf_ctx_last_pass(ctx: *mut rf_ctx, time_ns: *mut u64, n_launches: *mut u32) -> rf_status;
    pub fn rf_ctx_profile(ctx: *mut rf_ctx, enable: c_int) -> rf_status;
    pub fn rf_ctx_kernel_times(ctx: *mut rf_ctx, ns: *mut u64, launches: *mut u64) -> rf_status;
    pub fn rf_kernel_name(i: u32) -> *const c_char;
}
