// rf_device.cuh — device-side structures and exact-arithmetic helpers shared by the kernels.
//
// NUMERICS CONTRACT: every + - * / is one IEEE binary32 operation,
// round-to-nearest-even, never fused. This translation unit MUST be compiled with
//   -fmad=false -prec-div=true -prec-sqrt=true -ftz=false      (never --use_fast_math)
// Interpolation is by *sequential* accumulation (val = val + step), never closed form.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/retrofire_b200.h"

// Inline PTX (laneid aside) can be switched off: the SIMT emulation of tests/emu compiles the C++ equivalents instead.
#ifndef RF_SMEM_ASM
#define RF_SMEM_ASM 1
#endif

#define RF_TILE 32          // framebuffer tile: RF_TILE rows x RF_TILE columns, one warp owns one tile
#define RF_TILE_SHIFT 5
// TMA bulk copies (cp.async.bulk, SASS UBLKCP) move the depth tile between HBM and shared memory, one 128-byte row per lane:
// rows must then start on 16-byte boundaries, i.e. a pitch of 36 words instead of the conflict-free 33.
#ifndef RF_TMA_DEPTH
#define RF_TMA_DEPTH 0   // opt-in build variant: measured 2-8 % slower than the LDG/STG path at pitch 33 (profiles/r02_ab_tma_depth.txt)
#endif
#if RF_TMA_DEPTH
#define RF_TILE_PITCH 36    // smem row pitch in words: 144-byte rows (bank = (4 * row + col) % 32)
#else
#define RF_TILE_PITCH 33    // smem row pitch in words (bank = (row + col) % 32)
#endif
#define RF_MAX_ROWS (1u << 20)

// per-draw flag bits
#define RF_F_CULL_MASK 0x3u
#define RF_F_DTEST_SHIFT 2
#define RF_F_DTEST_MASK 0x3u
#define RF_F_CWRITE 0x10u
#define RF_F_DWRITE 0x20u
#define RF_F_RASTER_STATE ((RF_F_DTEST_MASK << RF_F_DTEST_SHIFT) | RF_F_CWRITE | RF_F_DWRITE)  // what the fragment stage reads
#define RF_F_SV 0x200u      // vertices are shared by enough primitives: k_vertex also stores their screen-space form
#define RF_F_BBOX 0x100u    // skip the draw when its bounding box is Hidden (scene.rs:81-87)
#define RF_F_DSORT_SHIFT 6   // Context::depth_sort (ctx.rs:39): RF_SORT_*
#define RF_F_DSORT_MASK 0x3u
#define RF_SORT_FRONT_TO_BACK 1u
#define RF_SORT_BACK_TO_FRONT 2u

// device error bits (PassStatus::error)
#define RF_ERRBIT_INDEX_OOB 0x1u
#define RF_ERRBIT_TARGET_OOB 0x2u
#define RF_ERRBIT_BIN_TOO_DEEP 0x4u
#define RF_ERRBIT_INTERNAL 0x10u
#define RF_ERRBIT_TEXEL_OOB 0x20u   // SamplerOnce indexed outside its texture (tex.rs:343-356 panics); raised by k_raster itself

struct DrawDesc {
  const float* verts;       // [n_verts][vstride]
  const uint32_t* indices;  // [n_prims][3]
  const uint32_t* tex;      // RGBA8 texels (device copy is always 4 B/texel), row-major, pitch = tex_w
  uint32_t vstride, n_verts, n_prims;
  uint32_t L, persp_mask, vs, fs;
  uint32_t prim_kind;       // RF_PRIM_TRIS / RF_PRIM_EDGES
  uint32_t target;          // index into the pass's TargetDesc table
  uint32_t flags;           // RF_F_*
  uint32_t tex_w, tex_h;
  float vs_u[RF_VS_UNIFORM_F32];
  float fs_u[RF_FS_UNIFORM_F32];
  float vp[12];             // rows 0..2 of the viewport matrix
  float bbox[6];            // BBox<Model> low, upp (RF_F_BBOX)
};

#define RF_MAX_PEERS 7
struct TargetDesc {
  uint32_t* color;  // uint32 containers, pitch = w
  float* depth;     // or nullptr
  uint32_t w, h, fmt;
  uint32_t tiles_x, tiles_y, tile_base;
  uint32_t band_y0, band_y1;
  // Sort-first over NVLink peer memory: every colour store of this GPU's band is replicated into the colour buffers of
  // the other GPUs (same layout), so each GPU ends the pass holding the whole frame without a separate gather.
  uint32_t n_peers;
  // First-touch clear (Frame::clear, front/src/lib.rs:103-120, recorded at the head of this pass): the rasteriser
  // initialises the tiles it touches itself (depth tile in shared memory from the clear value, colour rows stored
  // before the first fragment), k_clear_untouched fills the other tiles — every pixel is written once.
  uint32_t clear_flags;           // RF_CLEAR_COLOR | RF_CLEAR_DEPTH
  uint32_t clear_color, clear_zbits;
  // Lazy depth clear (RF_LAZY_DEPTH): three words per tile of the target — {lazy, zbits, slices done}. lazy != 0: every depth value
  // of the tile equals zbits and has NOT been written to `depth` (a first-touch clear marked the untouched tile instead of
  // filling it); the rasteriser that next touches the tile starts from zbits without a load and writes the tile back, after
  // which the tile is ordinary memory again; every host-visible read of the depth plane materialises first. nullptr: off.
  uint32_t* lazy;
  uint32_t* peer_color[RF_MAX_PEERS];
};
#define RF_LAZY_WORDS 3u
#define RF_CLEAR_COLOR 1u
#define RF_CLEAR_DEPTH 2u
#define RF_CLEAR_LAZY 4u    // with RF_CLEAR_DEPTH: the untouched tiles are marked in TargetDesc::lazy instead of being filled

struct DrawStats {
  unsigned long long prims_o, frags_i, frags_o;
  unsigned long long hidden;  // k_objects: the draw's bounding box is outside the frustum -> the draw does not happen
};

// zeroed before every pass; read back after it.
// Every allocation counter sits in its own 128-byte line: they are hit by one atomic per warp from
// thousands of warps, and atomics on one line serialise in a single L2 slice.
struct alignas(128) PaddedCounter {
  unsigned long long v;
  unsigned long long _pad[15];
  __host__ __device__ operator unsigned long long() const { return v; }
};
__device__ __forceinline__ unsigned long long atomicAdd(PaddedCounter* c, unsigned long long n) { return atomicAdd(&c->v, n); }

struct PassStatus {
  PaddedCounter stris_needed;    // screen triangles surviving clip + cull
  PaddedCounter spans_needed;    // span records the pass wants (one per scanline of every drawn triangle)
  PaddedCounter tris_needed;     // triangle records
  PaddedCounter entries_needed;  // bin entries (triangle x overlapped tile)
  PaddedCounter chunks_needed;   // walk chunks (<= 32 rows of one trapezoid half)
  PaddedCounter tall_needed;     // halves with more than one chunk
  PaddedCounter long_needed;     // spans that cross a tile-column boundary
  PaddedCounter ckpts_needed;    // checkpoint records for those spans
  PaddedCounter bins_needed;     // total size of all tile bins
  PaddedCounter small_needed;    // SMALL triangles: set up by k_assemble, binned there, walked by k_raster (no k_setup, no span records)
  uint32_t error;                // RF_ERRBIT_*
  uint32_t overflow;             // a capacity was exceeded: nothing was rasterised
  uint32_t n_work;               // non-empty tiles
  uint32_t n_work_big;           // tiles whose bin needs the large-smem sort
  uint32_t max_bin;
  uint32_t n_work_heavy;         // tiles with >= RF_HEAVY_BIN triangles
  uint32_t n_work_heaviest;      // tiles with >= RF_HEAVIEST_BIN triangles
  uint32_t _pad;
};

// Persistent across passes: the sequence number of the first pass that exceeded an arena capacity (RF_NO_POISON: none). That
// pass and every later one are no-ops until the host has grown the arenas and replayed them in order; EARLIER passes still
// in flight (the rasteriser of pass k runs next to the geometry stage of pass k + 1) are not disturbed.
#define RF_NO_POISON 0xFFFFFFFFu
struct CtxStatus {
  uint32_t poison;
};

struct ClearDesc {
  uint32_t* ptr;
  unsigned long long n;
  uint32_t value;
  uint32_t n_lazy;   // depth plane of a target with lazy tiles: the words of `lazy` to zero (the plane is written as a whole)
  uint32_t* lazy;
};

struct PassParams {
  const DrawDesc* draws;
  const uint32_t* vbase;  // [n_draws+1] prefix of n_verts
  const uint32_t* pbase;  // [n_draws+1] prefix of n_prims
  const TargetDesc* targets;
  uint32_t n_draws, n_targets, NV, NP, n_tiles;
  uint32_t seq;           // pass sequence number (see CtxStatus)
  uint32_t use_sv;        // every draw of the pass has RF_F_SV: k_vertex stores screen-space vertices, k_assemble<LT, true> reads them
  uint32_t any_bbox;      // some draw of the pass carries RF_F_BBOX (k_objects ran)
  uint32_t tiles_per_target;  // every target of the pass has this many tiles (tile / it = target index), or 0 when they differ
  uint32_t fused_clear;   // the first-touch clear of the tiles WITHOUT bin entries is done by k_raster's warps (no k_clear_untouched launch)
  uint32_t verts_per_draw, prims_per_draw;  // likewise for the draws of a frame batch (one mesh, many frames): index / it = draw, or 0
  float* cv;              // clip verts [NV][CVS]
  float* sv;              // screen verts [NV][SVS]: to_screen of every vertex inside the frustum, plus its outcode
  uint32_t* stris;        // [cap_stris][QW]   compacted screen triangles (k_assemble -> k_setup)
  uint32_t cap_stris;
  uint32_t* smalls;       // [cap_smalls][SmallRec<LT>::W]  SMALL triangles: both trapezoid-half setups, written by k_assemble, walked by k_raster
  uint32_t cap_smalls;
  uint32_t* sdepth;       // [cap_stris] total-order bits of Render::depth, or null when no draw of the pass is depth-sorted
  uint32_t* spans;        // [cap_spans][SW]   one per scanline, contiguous per triangle
  uint32_t* tris;         // [cap_tris][TW]    per drawn triangle: key, draw, rows, dv/dx of both halves
  uint4* entries;         // [cap_entries]     {tile, key, tri, 0}: one per (triangle, overlapped tile)
  unsigned long long* bins;  // [cap_entries]  per-tile lists of (key << 32 | tri), sorted by k_bin_sort
  unsigned long long* bins2; // [cap_entries]  merge scratch for bins deeper than one shared-memory run
  uint4* chunks;          // [cap_chunks]      {tri*2+half, chunk | rows << 16, first span, draw | target << 16}
  uint32_t* talllist;     // [cap_tall]        tri*2+half of halves with more than one chunk
  uint32_t* ecks;         // [cap_chunks][EW]  edge state at the start of a chunk (indexed by chunk position; chunk 0 unused)
  uint32_t cap_chunks, cap_tall, cap_ecks;
  uint2* longlist;        // [cap_long]        {span index, tri*2+half} of spans crossing a tile-column boundary
  uint32_t* ckpts;        // [cap_ckpts][KW]   varyings of such spans at each later tile-column start
  uint32_t cap_spans, cap_tris, cap_entries, cap_long, cap_ckpts;
  uint32_t* tile_cnt;     // [n_tiles]
  uint32_t* tile_off;     // [n_tiles]
  uint32_t* tile_fill;    // [n_tiles]
  uint32_t* worklist;     // [n_tiles] non-empty tiles
  uint32_t* worklist_big; // [n_tiles]
  uint32_t* worklist_heavy; // [(RF_SLICES + 1) * n_tiles] rasterised first: slice tasks of the heaviest tiles, then the heavy tiles
  uint32_t* cursors;      // [1] raster work cursor, [2] tile cursor of the fused first-touch clear
  DrawStats* dstats;      // [n_draws]
  PassStatus* status;
  CtxStatus* cstatus;
};

__device__ __forceinline__ bool rf_poisoned(const PassParams& P) { return P.cstatus->poison <= P.seq; }
__device__ __forceinline__ void rf_overflow(const PassParams& P) { P.status->overflow = 1; atomicMin(&P.cstatus->poison, P.seq); }

#define RF_NO_CKPT 0xFFFFFFFFu
#define RF_STRI_LINE 0x80000000u  // flag in the draw word of a screen-triangle record: the record is an Edge (2 vertices)
#define RF_NO_TILE 0xFFFFFFFFu
// Bin entries name either a triangle record written by k_setup, or — flag set — a SMALL-triangle record of k_assemble
// (few scanlines, narrow, inside the target) whose scanlines the rasteriser walks itself.
#define RF_BIN_SMALL 0x80000000u
#ifndef RF_SMALL_ROWS
#define RF_SMALL_ROWS 32u      // most scanlines of a SMALL triangle (0 switches the class off)
#endif
#ifndef RF_SMALL_WIDTH
#define RF_SMALL_WIDTH 64.0f   // widest bounding box (pixels) of a SMALL triangle
#endif

// record strides in 32-bit words, as a function of the compile-time lane count LT
template <int LT> struct Rec {
  static constexpr int CVS = (5 + LT + 3) & ~3;        // clip vert: pos4, oc, attr[LT]            (16 B aligned)
  static constexpr int SVS = (4 + LT + 3) & ~3;        // screen vert: x, y, z, attr[LT], oc       (16 B aligned)
  static constexpr int SW = (2 + 1 + LT + 1) & ~1;     // span: X0|n<<16, ckpt, z, attr[LT]         (8 B aligned)
  static constexpr int HS = (3 * LT + 8 + 3) & ~3;      // half setup: dv[1+LT], L[2+LT], dl[2+LT], R, dr, y    (16 B aligned)
  static constexpr int TW = 8 + 2 * HS;                // tri: 8 header words + two half setups (see TriRec)
  static constexpr int EW = (2 + LT + 1 + 1) & ~1;     // edge checkpoint: L[2+LT], R                        (8 B aligned)
  static constexpr int KW = (1 + LT + 1) & ~1;         // checkpoint: z, attr[LT]                   (8 B aligned)
  static constexpr int QW = (2 + 3 * (3 + LT) + 3) & ~3;  // screen triangle: key, draw, 3 x (x,y,z,attr[LT]) (16 B aligned)
};

// SMALL triangle record (k_assemble -> k_raster): [0] key  [1] draw  [2] first row Y0  [3] rows of the upper half | lower half << 16,
// then per trapezoid half (HW words, 16-byte aligned): L[2+LT] (x, z, attr at the first row centre), dl[2+LT] (per-row step),
// R, dr, dv/dx[1+LT] (z, attr) — everything raster.rs:248-302 precomputes, minus the unused y lane. 144 bytes at LT = 3.
template <int LT> struct SmallRec {
  static constexpr int NV = 1 + LT, NL = 2 + LT;
  static constexpr int O_L = 0, O_DL = NL, O_R = 2 * NL, O_DR = 2 * NL + 1, O_DV = 2 * NL + 2;
  static constexpr int HW = (NV + 2 * NL + 2 + 3) & ~3;
  static constexpr int W = 4 + 2 * HW;
};

// ---- Rust `as` casts (saturating, NaN -> 0). PTX cvt.rzi.{u32,s32}.f32 clamps and maps NaN to 0.
__device__ __forceinline__ uint32_t sat_u32(float f) { return __float2uint_rz(f); }
__device__ __forceinline__ int32_t sat_i32(float f) { return __float2int_rz(f); }
// `f as u8` (color.rs:253-263): truncate, clamp to 0..255, NaN -> 0 — one F2I.U8 instead of F2I.U32 + IMNMX
__device__ __forceinline__ uint32_t sat_u8(float f) {
#if RF_SMEM_ASM
  uint32_t r;
  asm("cvt.rzi.u8.f32 %0, %1;" : "=r"(r) : "f"(f));
  return r;
#else
  return min(__float2uint_rz(f), 255u);
#endif
}

// math/vec.rs:231-238: left fold from 0.0
__device__ __forceinline__ float dot4(float a0, float a1, float a2, float a3, float b0, float b1, float b2, float b3) {
  float acc = 0.0f;
  acc = acc + a0 * b0;
  acc = acc + a1 * b1;
  acc = acc + a2 * b2;
  acc = acc + a3 * b3;
  return acc;
}
__device__ __forceinline__ float dot4p(const float* __restrict__ m, float b0, float b1, float b2, float b3) {
  return dot4(m[0], m[1], m[2], m[3], b0, b1, b2, b3);
}

// render/clip.rs:215-222,240-242. Plane vectors are (n, -1); dot keeps the zero terms, as the reference does.
__device__ __forceinline__ float plane_dist(int p, float x, float y, float z, float w) {
  switch (p) {
    case 0: return dot4(0.0f, 0.0f, -1.0f, -1.0f, x, y, z, w);
    case 1: return dot4(0.0f, 0.0f, 1.0f, -1.0f, x, y, z, w);
    case 2: return dot4(-1.0f, 0.0f, 0.0f, -1.0f, x, y, z, w);
    case 3: return dot4(1.0f, 0.0f, 0.0f, -1.0f, x, y, z, w);
    case 4: return dot4(0.0f, -1.0f, 0.0f, -1.0f, x, y, z, w);
    default: return dot4(0.0f, 1.0f, 0.0f, -1.0f, x, y, z, w);
  }
}
__device__ __forceinline__ uint32_t outcode(float x, float y, float z, float w) {
  uint32_t oc = 0;
#pragma unroll
  for (int p = 0; p < 6; p++) oc |= (plane_dist(p, x, y, z, w) > 0.0f) ? (1u << p) : 0u;
  return oc;
}

__device__ __forceinline__ float lerpf(float a, float b, float t) { return a + (b - a) * t; }  // math.rs:196-198
__device__ __forceinline__ float round_up_to_half(float x) { return floorf(x + 0.5f) + 0.5f; }  // raster.rs:304-307

// ZDiv for f32 (math/vary.rs:135-140): a / z, correctly rounded. An exactly-zero numerator (a component of an axis-aligned
// normal, a uv on a texture edge) would send ptxas' div.rn sequence into its ~50-instruction slow path on every fragment
// (FCHK flags zero operands); the quotient is then a zero with the sign of a XOR z whenever z is neither NaN nor zero.
__device__ __forceinline__ float zdiv(float a, float z) {
  if (a == 0.0f && z == z && z != 0.0f) return __uint_as_float((__float_as_uint(a) ^ __float_as_uint(z)) & 0x80000000u);
  return a / z;
}

// f32::total_cmp key (stable sort of the three vertices by y, raster.rs:191)
__device__ __forceinline__ int32_t total_key(float f) {
  int32_t b = __float_as_int(f);
  return b ^ (int32_t)(((uint32_t)(b >> 31)) >> 1);
}

// util/pixfmt.rs:45-142: Color4 -> uint32 container
__device__ __forceinline__ uint32_t pack_pixel(uint32_t fmt, uint32_t r, uint32_t g, uint32_t b, uint32_t a) {
  switch (fmt) {
    case RF_FMT_RGBA8888: return r | g << 8 | b << 16 | a << 24;
    case RF_FMT_XRGB8888: return r << 16 | g << 8 | b;
    case RF_FMT_ARGB8888: return a | r << 8 | g << 16 | b << 24;
    case RF_FMT_BGRA8888: return b | g << 8 | r << 16 | a << 24;
    case RF_FMT_RGB888: return r | g << 8 | b << 16;
    case RF_FMT_RGB565: return ((r >> 3) & 0x1Fu) << 11 | ((g >> 2) & 0x3Fu) << 5 | ((b >> 3) & 0x1Fu);
    default: return (r >> 4) << 12 | (g >> 4) << 8 | (b >> 4) << 4 | (a >> 4);  // RGBA4444
  }
}

// The four 8888 layouts and RGB888 are byte permutations of r | g << 8 | b << 16 | a << 24: one PRMT with a per-target
// selector (0: not such a format, use pack_pixel). Index 4 selects a zero byte from the second PRMT operand.
__host__ __device__ __forceinline__ uint32_t pack_selector(uint32_t fmt) {
  switch (fmt) {
    case RF_FMT_RGBA8888: return 0x3210u;
    case RF_FMT_XRGB8888: return 0x4012u;
    case RF_FMT_ARGB8888: return 0x2103u;
    case RF_FMT_BGRA8888: return 0x3012u;
    case RF_FMT_RGB888: return 0x4210u;
    default: return 0u;
  }
}
__device__ __forceinline__ uint32_t pack_pixel_sel(uint32_t sel, uint32_t fmt, uint32_t r, uint32_t g, uint32_t b, uint32_t a) {
  if (sel) return __byte_perm(r | g << 8 | b << 16 | a << 24, 0u, sel);
  return pack_pixel(fmt, r, g, b, a);
}

// binary search: largest d with base[d] <= i   (base has n+1 entries, base[0] = 0). `per` != 0: every draw has `per`
// items (a frame batch), so the draw is one division instead of log2(n) dependent loads.
__device__ __forceinline__ uint32_t find_draw(const uint32_t* __restrict__ base, uint32_t n, uint32_t i, uint32_t per = 0) {
  if (per) return i / per;
  uint32_t lo = 0, hi = n;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (__ldg(base + mid) <= i) lo = mid; else hi = mid;
  }
  return lo;
}

// Lane index, read ONCE per kernel. The asm is volatile on purpose: otherwise the compiler prefers to
// re-read the special register (S2R, a slow XU-pipe instruction) inside hot loops instead of keeping
// the value in a register; kernels pass `lane` (and the lanes-below mask `lt`) to the helpers below.
__device__ __forceinline__ uint32_t lane_id() {
  uint32_t l;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
  return l;
}

// inclusive warp scan
__device__ __forceinline__ uint32_t warp_scan_incl(uint32_t v, uint32_t lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, o);
    if ((int)lane >= o) v += t;
  }
  return v;
}

// Warp-aggregated increment usable from divergent code: the currently converged lanes elect a
// leader that performs one atomic for the group; returns this lane's slot.
__device__ __forceinline__ unsigned long long agg_atomic_inc(PaddedCounter* ctr, uint32_t lane) {
  const uint32_t mask = __activemask();
  const int leader = __ffs(mask) - 1;
  unsigned long long base = 0;
  if ((int)lane == leader) base = atomicAdd(ctr, (unsigned long long)__popc(mask));
  base = __shfl_sync(mask, base, leader);
  return base + __popc(mask & ((1u << lane) - 1u));
}
