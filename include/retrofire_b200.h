/*
 * retrofire_b200.h — C ABI of the B200-native replacement for retrofire's render() hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types. A Rust shim
 * crate (rust/retrofire-b200-sys) binds exactly these symbols and
 * re-exposes retrofire-core's `render()` / `Batch` / `Context` / `Stats` / `Buf2` /
 * `Texture` API on top of them. Citations are file:line in the reference tree
 * (jdahlstrom/retrofire v0.4.0, `core/src/...`).
 *
 * Threading: an rf_ctx is single-owner (the reference's Context is !Sync,
 * render/ctx.rs:62-64). All calls on one ctx must come from one thread at a time.
 *
 * Execution model: rf_render() *queues* a draw on the ctx; queued draws are executed in
 * submission order, as one device pass, at rf_flush(), at any rf_target_download_*(),
 * rf_ctx_stats() or rf_sync(). The reference's "pixels are in the buffer when render()
 * returns" is preserved at the observable boundary (download), and a caller that wants
 * the per-call behaviour passes a non-NULL `stats_out` to rf_render(), which flushes and
 * synchronises before returning.
 *
 * There is NO CPU fallback. Every entry point fails with RF_E_CUDA if no sm_100 device
 * is usable.
 */
#ifndef RETROFIRE_B200_H
#define RETROFIRE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RF_ABI_VERSION 2   /* 2: rf_mesh_create takes the primitive kind; RF_N_KERNELS 12 */
#define RF_MAX_ATTR_LANES 8   /* varying lanes besides position (x,y,z) */
#define RF_VS_UNIFORM_F32 32  /* e.g. two row-major 4x4 matrices */
#define RF_FS_UNIFORM_F32 8

typedef struct rf_ctx rf_ctx;         /* one per GPU: stream, arenas, queued draws, Stats     */
typedef struct rf_target rf_target;   /* device-resident Framebuf / Colorbuf / Buf2<Color4>   */
typedef struct rf_texture rf_texture; /* device-resident Texture<Buf2<Color3|Color4>>         */
typedef struct rf_mesh rf_mesh;       /* optional persistent vertex+index data                */

/* The reference reports these conditions by panicking (SURVEY §8b "Errors"); the shim
 * converts a non-zero status back into panic!(). */
typedef enum rf_status {
  RF_OK = 0,
  RF_E_INVALID = 1,            /* NULL/ill-formed argument                                    */
  RF_E_INDEX_OOB = 2,          /* vertex index >= n_verts          (render/prim.rs:17-19)      */
  RF_E_TARGET_OOB = 3,         /* scanline outside the target      (render/target.rs:148,173)  */
  RF_E_BAD_TEXTURE = 4,        /* non-POT texture with RepeatPot   (render/tex.rs:230-231)     */
  RF_E_UNSUPPORTED_SHADER = 5, /* vs/fs not in the catalogue, or lanes do not fit it           */
  RF_E_CUDA = 6,
  RF_E_NCCL = 7,
  RF_E_NOMEM = 8,
  RF_E_UNSUPPORTED = 9         /* an option outside the path (e.g. async download of a non-32-bit format) */
} rf_status;

/* Vertex-shader catalogue. Input vertex = [x,y,z,a0..]; output = clip pos
 * + `n_attr_lanes` varying lanes. vs_uniform layout is given per entry (row-major 4x4). */
typedef enum rf_vs_id {
  RF_VS_MVP = 0,           /* u[0..16]=mvp. pos'=mvp·[p,1]; attribs pass through.
                              demos crates.rs:32,39-41; core/tests/rendering.rs:27-29; hello_tri.rs:22-26 */
  RF_VS_MVP_LINEARIZE = 1, /* as MVP, attrib lanes -> powf(c, 2.2)   hello_tri.rs:13-17 (fp)  */
  RF_VS_SOLIDS = 2,        /* u[0..16]=mvp, u[16..32]=spin. in: normal3, out: Color3f
                              demos/src/bin/solids.rs:70-79                                   */
  RF_VS_SPRITE = 3         /* u[0..16]=modelview, u[16..32]=proj. in/out: vec2
                              demos/src/bin/sprites.rs:40-45                                  */
} rf_vs_id;

/* Fragment-shader catalogue. */
typedef enum rf_fs_id {
  RF_FS_COLOR3F = 0,        /* 3 colour lanes -> (256*c) as u8, a=255   solids.rs:81-83       */
  RF_FS_COLOR3F_SRGB = 1,   /* powf(c,1/2.2) first                      hello_tri.rs:18 (fp)  */
  RF_FS_COLOR4F = 2,        /* 4 colour lanes                           render/debug.rs:34-38 */
  RF_FS_CHECKER = 3,        /* 2 lanes: (x>.5)^(y>.5) ? .8 : .1 gray    crates.rs:33-36       */
  RF_FS_TEX_CLAMP_LIT = 4,  /* 5 lanes n3+uv2, fs_uniform[0..3]=light   crates.rs:42-47       */
  RF_FS_TEX_CLAMP = 5,      /* 2 lanes uv, SamplerClamp                 tests/rendering.rs:30 */
  RF_FS_TEX_REPEAT_POT = 6, /* 2 lanes uv, SamplerRepeatPot             benches/fill.rs:74-91 */
  RF_FS_SPRITE_DISC = 7,    /* 2 lanes; d2<1 ? 1-d2*(.25,.5,1) : discard sprites.rs:46-52     */
  RF_FS_NORMAL_VIS = 8,     /* 3 lanes; n/2+0.5                         curses.rs:53-56       */
  RF_FS_TEX_ONCE = 9        /* 2 lanes uv, SamplerOnce: texel (w*u as u32, h*v as u32), unchecked; a coordinate outside the
                               texture panics in the reference (tex.rs:313-357) = RF_E_BAD_TEXTURE here */
} rf_fs_id;

/* Colour element of the target (util/pixfmt.rs:20-142, render/target.rs:99-136).
 * On the device every pixel is one uint32 container; host layouts are converted in
 * rf_target_upload_color/rf_target_download_color. */
typedef enum rf_color_fmt {
  RF_FMT_RGBA8888 = 0, /* [u8;4] r,g,b,a   — Buf2<Color4>, Colorbuf<_,Rgba8888>; host 4 B/px   */
  RF_FMT_XRGB8888 = 1, /* u32 0x00RRGGBB   — Colorbuf<u32,Xrgb8888>;             host 4 B/px   */
  RF_FMT_ARGB8888 = 2, /* [u8;4] a,r,g,b                                          host 4 B/px  */
  RF_FMT_BGRA8888 = 3, /* [u8;4] b,g,r,a                                          host 4 B/px  */
  RF_FMT_RGB888 = 4,   /* [u8;3] r,g,b     — Buf2<Color3>;                        host 3 B/px  */
  RF_FMT_RGB565 = 5,   /* [u8;2] native-endian u16                                host 2 B/px  */
  RF_FMT_RGBA4444 = 6  /* [u8;2] native-endian u16                                host 2 B/px  */
} rf_color_fmt;

typedef enum rf_texel_fmt {
  RF_TEXEL_RGB888 = 0,  /* Color3, 3 B/texel */
  RF_TEXEL_RGBA8888 = 1 /* Color4, 4 B/texel */
} rf_texel_fmt;

enum { RF_PRIM_TRIS = 0, RF_PRIM_EDGES = 1 };
enum { RF_SORT_NONE = 0, RF_SORT_FRONT_TO_BACK = 1, RF_SORT_BACK_TO_FRONT = 2 }; /* ctx.rs:39,67-72 */
enum { RF_CULL_NONE = 0, RF_CULL_BACK = 1, RF_CULL_FRONT = 2 };                /* ctx.rs:74-78 */
enum { RF_DEPTH_NONE = 0, RF_DEPTH_LESS = 1, RF_DEPTH_EQUAL = 2, RF_DEPTH_GREATER = 3 }; /* ctx.rs:42-48,86-89 */

/* One render() call: render.rs:134-147. Geometry comes either from host pointers
 * (copied during the call, as the reference borrows them only for the call) or from a
 * persistent rf_mesh (then verts/indices must be NULL). */
typedef struct rf_draw {
  const uint32_t* indices;  /* 3 per primitive (Tri<usize>, geom/prim.rs:31-33), 2 for RF_PRIM_EDGES */
  uint32_t n_prims;
  const float* verts;       /* n_verts records of vert_stride_f32 floats: [x,y,z,a0..a(L-1)]  */
  uint32_t n_verts;
  uint32_t vert_stride_f32; /* >= 3 + n_attr_lanes                                            */
  const rf_mesh* mesh;      /* or NULL                                                        */
  uint32_t n_attr_lanes;    /* L, 0..RF_MAX_ATTR_LANES                                        */
  uint32_t persp_mask;      /* bit i set: lane i is divided in z_div (f32/Vector/Point lanes;
                               clear for Color lanes)   math/vary.rs:10-15, math/color.rs:628 */
  uint32_t vs;              /* rf_vs_id                                                       */
  uint32_t fs;              /* rf_fs_id                                                       */
  float vs_uniform[RF_VS_UNIFORM_F32];
  float fs_uniform[RF_FS_UNIFORM_F32];
  const rf_texture* texture; /* for the RF_FS_TEX_* entries, else NULL                        */
  float viewport[16];       /* Mat4<Ndc,Screen>, row-major   math/mat.rs:1304-1315            */
  /* Context fields (render/ctx.rs:11-65) consumed by the path */
  uint8_t face_cull;        /* RF_CULL_*    default BACK                                      */
  uint8_t depth_test;       /* RF_DEPTH_*   default LESS                                      */
  uint8_t color_write;      /* default 1                                                      */
  uint8_t depth_write;      /* default 1                                                      */
  uint8_t depth_sort;       /* RF_SORT_*    default NONE. Equal depths keep primitive order (the reference's
                               sort_unstable_by leaves their order unspecified)  render.rs:180-182,209-219 */
  uint8_t prim_kind;        /* RF_PRIM_TRIS: 3 indices per primitive (Tri<usize>); RF_PRIM_EDGES: 2 indices per
                               primitive (Edge<usize>: render/prim.rs:41-60, clip.rs:311-348, raster.rs:122-177) */
  uint8_t bbox_cull;        /* 1: scene-style object culling on the device (render/scene.rs:81-87, crates.rs:100-122):
                               when BBox::visibility(vs_uniform[0..16]) of `bbox` is Hidden the draw is skipped as if
                               render() had not been called (no Stats but objs.i). Needs a VS whose u[0..16] is the
                               model-to-projection matrix (every catalogue VS except RF_VS_SPRITE). */
  uint8_t _pad[1];
  float bbox[6];            /* BBox<Model>: low x,y,z then upp x,y,z   scene.rs:22 */
} rf_draw;

/* render/stats.rs:16-40. time_ns is device time of the pass(es) (CUDA events). */
typedef struct rf_stats {
  uint64_t calls;
  uint64_t prims_i, prims_o;
  uint64_t verts_i, verts_o;
  uint64_t frags_i, frags_o;
  uint64_t time_ns;
  uint64_t objs_i, objs_o;  /* draws submitted with bbox_cull / those of them not Hidden (crates.rs:101,131) */
} rf_stats;

/* ---- context ------------------------------------------------------------------------- */
uint32_t rf_abi_version(void);
/* `stream` is a cudaStream_t to run on (e.g. torch's current stream) or NULL to create one. */
rf_status rf_ctx_create(int device, void* stream, rf_ctx** out);
void rf_ctx_destroy(rf_ctx* ctx);
const char* rf_last_error(const rf_ctx* ctx);
/* Sort-first sharding: this ctx rasterises only framebuffer rows y with
 * y0 <= y < y1 of every target (geometry is replicated). Default: all rows. frags.i/o then
 * count only the owned rows. */
rf_status rf_ctx_set_row_band(rf_ctx* ctx, uint32_t y0, uint32_t y1);

/* ---- targets: Buf2 / Colorbuf / Framebuf (render/target.rs:35-136) ---------------------- */
rf_status rf_target_create(rf_ctx* ctx, uint32_t w, uint32_t h, uint32_t color_fmt,
                           int has_depth, rf_target** out); /* zero-initialised, util/buf.rs:155-161 */
void rf_target_destroy(rf_target* t);
/* Frame::clear (front/src/lib.rs:103-120): rgba = Color4 to convert with the target's
 * format, or NULL to keep colour; depth_recip = value to fill (already 1/depth_clear) or NULL. */
rf_status rf_target_clear(rf_ctx* ctx, rf_target* t, const uint8_t* rgba, const float* depth_recip);
/* Host <-> device. stride in ELEMENTS of the host buffer (util/buf.rs:437-439). Downloads
 * flush queued draws and synchronise. */
rf_status rf_target_upload_color(rf_ctx* ctx, rf_target* t, const void* host, size_t stride_elems);
rf_status rf_target_download_color(rf_ctx* ctx, rf_target* t, void* host, size_t stride_elems);
rf_status rf_target_upload_depth(rf_ctx* ctx, rf_target* t, const float* host, size_t stride_elems);
rf_status rf_target_download_depth(rf_ctx* ctx, rf_target* t, float* host, size_t stride_elems);
/* Pinned host memory for Buf2 storage, so that uploads/downloads DMA directly (no staging copy). */
/* Opt-in: vertex/index arrays that live in page-locked memory (rf_host_alloc) are read by DMA AFTER rf_render has
 * returned; the caller must leave them unchanged until the next rf_flush or rf_sync returns (the same contract as
 * rf_target_download_color_async in the other direction). Default off: rf_render returns only when its arrays
 * have been read, like the borrows of render(). Pageable arrays are always copied during the call. */
rf_status rf_ctx_set_geometry_async(rf_ctx* ctx, int on);
rf_status rf_host_alloc(size_t bytes, void** out);
void rf_host_free(void* p);
/* Asynchronous download (4-byte formats): ordered after the queued draws on the ctx stream; the
 * pixels are valid after the next rf_sync(). A replayed (arena-overflow) pass re-issues nothing:
 * callers that use this must have warmed the arenas up, or use the synchronous download. */
rf_status rf_target_download_color_async(rf_ctx* ctx, rf_target* t, void* host, size_t stride_elems);
/* Device pointers of the uint32 colour containers / float depth (row-major, stride = w),
 * for NCCL gathers done by the host layer. The depth plane is scratch state of the rasteriser: a
 * Frame::clear recorded with draws only MARKS the 32x32 tiles the pass does not touch ("lazy depth
 * clear"), and every read through this API sees the cleared values. rf_target_depth_devptr flushes
 * the queued draws, writes the marked tiles out on the ctx stream and keeps the plane materialised
 * from then on, so the pointer can be used like any device buffer ordered on that stream. */
void* rf_target_color_devptr(rf_target* t);
void* rf_target_depth_devptr(rf_target* t);

/* ---- textures (render/tex.rs:33-37) ------------------------------------------------------ */
rf_status rf_texture_create(rf_ctx* ctx, uint32_t w, uint32_t h, uint32_t texel_fmt,
                            const void* data, size_t stride_elems, rf_texture** out);
void rf_texture_destroy(rf_texture* t);

/* ---- persistent geometry (Batch clones prims/verts per call, batch.rs:62-84; this avoids it).
 * prim_kind: RF_PRIM_TRIS (indices = n_prims x 3) or RF_PRIM_EDGES (n_prims x 2); a draw that names the mesh
 * with another rf_draw.prim_kind is rejected with RF_E_INVALID. */
rf_status rf_mesh_create(rf_ctx* ctx, const float* verts, uint32_t n_verts,
                         uint32_t vert_stride_f32, const uint32_t* indices, uint32_t n_prims,
                         uint32_t prim_kind, rf_mesh** out);
void rf_mesh_destroy(rf_mesh* m);

/* ---- the hot path ------------------------------------------------------------------------ */
/* render() (render.rs:134-207). Queues the draw; if stats_out != NULL, flushes, waits and
 * returns this call's Stats (calls=1). Stats are always accumulated into the ctx
 * (`*ctx.stats.borrow_mut() += stats`, render.rs:206). */
rf_status rf_render(rf_ctx* ctx, rf_target* target, const rf_draw* draw, rf_stats* stats_out);
/* Frame batches: draw i goes to targets[i] with vs_uniform taken from
 * vs_uniforms + i*RF_VS_UNIFORM_F32; everything else from `draw`. Queued like rf_render. */
/* A frame's list of render() calls into one target in ONE call (the loop of crates.rs:114-131,
 * front/src/minifb.rs frame callbacks): same as n_draws rf_render(…, NULL) calls, in order. */
rf_status rf_render_many(rf_ctx* ctx, rf_target* target, const rf_draw* draws, uint32_t n_draws);
rf_status rf_render_frames(rf_ctx* ctx, rf_target* const* targets, uint32_t n_frames,
                           const rf_draw* draw, const float* vs_uniforms);
/* ---- sort-first over NVLink peer memory --------------------------------------
 * With peers attached, every colour store of this GPU's row band (rf_ctx_set_row_band) is also
 * stored into the same pixel of the other GPUs' colour buffers (P2P over NVLink), inside the
 * rasteriser: after rf_sync on every GPU each of them holds the whole frame, with no separate
 * gather. Passes that draw into such a target run two cross-GPU barriers (after the clear, after
 * rasterisation), so EVERY GPU must submit the same sequence of clears/draws/flushes for it, and
 * rf_target_clear clears all rows of the colour buffer (depth: the band). Peers are named either
 * by CUDA IPC handles (one process per GPU) or by raw device pointers (one process, several ctxs).
 * Tables hold `world` entries in rank order; the entry of `rank` itself is ignored. In a target's
 * table an all-zero handle / NULL pointer means "do not push into that rank": with only a root
 * rank's entry set on the others (and an empty table on the root) the frame is gathered to root.
 * If rf_ctx_replays grew on ANY GPU during a frame (a pass was re-run after arena growth), peers
 * may have read the frame before the re-run's stores: render that frame again. */
#define RF_IPC_HANDLE_BYTES 64
#define RF_MAX_PEER_GPUS 8
rf_status rf_ctx_peer_export(rf_ctx* ctx, uint8_t* ipc_handle_out /* 64 B or NULL */, void** devptr_out /* or NULL */);
rf_status rf_ctx_peer_attach(rf_ctx* ctx, uint32_t world, uint32_t rank, const uint8_t* ipc_handles, void* const* devptrs);
rf_status rf_target_peer_export(rf_ctx* ctx, rf_target* t, uint8_t* ipc_handle_out, void** devptr_out);
rf_status rf_target_peer_attach(rf_ctx* ctx, rf_target* t, uint32_t world, uint32_t rank, const uint8_t* ipc_handles, void* const* devptrs);
rf_status rf_ctx_replays(rf_ctx* ctx, uint64_t* out); /* passes re-launched after an arena overflow so far */

rf_status rf_flush(rf_ctx* ctx); /* execute queued draws (asynchronous on the stream)       */
rf_status rf_sync(rf_ctx* ctx);  /* flush + wait; reports deferred device-side errors        */
/* Accumulated Stats of the ctx (flushes + waits); reset=1 zeroes them afterwards. */
rf_status rf_ctx_stats(rf_ctx* ctx, rf_stats* out, int reset);
/* Device time in ns of the most recent executed pass, and the number of kernels it launched. */
rf_status rf_ctx_last_pass(rf_ctx* ctx, uint64_t* time_ns, uint32_t* n_launches);


/* ---- measurement (Stats::start/finish analogue, render/stats.rs:57-78, per kernel) ----------- */
#define RF_N_KERNELS 12 /* kernels launched by one pass, in order; see rf_kernel_name */
/* level 0: off. 1: CUDA events around k_raster only (the pass keeps its two-stream overlap).
 * 2: events between all pass kernels, which are then serialised on the ctx stream. Resets the sums. */
rf_status rf_ctx_profile(rf_ctx* ctx, int level);
/* Device time (ns) and launch count per kernel accumulated since the last call; resets them. */
rf_status rf_ctx_kernel_times(rf_ctx* ctx, uint64_t* ns /*[RF_N_KERNELS]*/, uint64_t* launches /*[RF_N_KERNELS]*/);
const char* rf_kernel_name(uint32_t i);

#ifdef __cplusplus
}
#endif
#endif /* RETROFIRE_B200_H */
