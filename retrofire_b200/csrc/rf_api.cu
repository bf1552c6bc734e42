// rf_api.cu — C ABI (include/retrofire_b200.h) and pass orchestration for the sm_100a kernels.
//
// A "pass" is every draw queued between two flush points, executed in submission order by
// one chain of kernels:
//   [k_clear_multi] -> k_vertex -> k_assemble -> k_setup -> k_edge_ckpt -> k_walk -> k_bin_alloc -> k_bin_scatter -> k_ckpt -> k_bin_sort_* -> k_raster
//   (+ the depth-sort ordering step of rf_order.cuh between k_assemble and k_setup for passes with Context::depth_sort)
// Passes are launched asynchronously on the ctx stream and validated lazily (capacity overflow
// or device-detected errors) at the next synchronisation point; an overflowing pass poisons
// the ctx on the device so that later passes become no-ops until the host has grown the
// arenas and replayed them in order.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -prec-div=true
//        -prec-sqrt=true -ftz=false -shared -Xcompiler -fPIC   (see __graft_entry__.build()).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "rf_device.cuh"
#include "rf_geometry.cuh"
#include "rf_raster.cuh"
#include "rf_order.cuh"
#include "rf_peer.cuh"


namespace {

constexpr size_t kTileArrays = 5 + RF_SLICES + 1;  // tile_cnt, tile_off, tile_fill, worklist, worklist_big, worklist_heavy[(RF_SLICES + 1) * n_tiles]
constexpr int kSlots = 3;          // passes that may be in flight before the oldest is validated
constexpr uint32_t kMaxTargetDim = 32768;

struct PinnedBuf {
  uint8_t* p = nullptr;
  size_t cap = 0;
  bool reserve(size_t n) {
    if (n <= cap) return true;
    size_t nc = std::max(n, cap * 2);
    uint8_t* q = nullptr;
    if (cudaHostAlloc(&q, nc, cudaHostAllocDefault) != cudaSuccess) return false;
    if (p) { std::memcpy(q, p, cap); cudaFreeHost(p); }
    p = q; cap = nc;
    return true;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  bool reserve(size_t n) {  // contents are NOT preserved
    if (n <= cap) return true;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    if (cudaMalloc(&p, n) != cudaSuccess) return false;
    cap = n;
    return true;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct QueuedDraw {
  DrawDesc desc;        // verts/indices/tex resolved to device pointers at launch
  rf_target* target;
  size_t verts_off;     // into geometry staging, or SIZE_MAX when the draw uses an rf_mesh
  size_t idx_off;
  bool direct;          // offsets are into the slot's d_direct buffer (page-locked caller memory, DMA'd during rf_render)
};

struct HostStatus {     // pinned, one per slot
  PassStatus status;
};

struct QueuedClear {  // Frame::clear recorded at the head of a pass (replayed with it)
  rf_target* target;
  bool has_color, has_depth;
  uint32_t color, zbits;
};

struct PassSlot {
  bool in_flight = false;
  std::vector<QueuedClear> clears;
  std::vector<QueuedDraw> draws;
  uint64_t prims_queued = 0;  // sum of draws[i].desc.n_prims (a pass addresses prims with 29 bits)
  std::vector<rf_target*> targets;
  PinnedBuf geom;       // pinned copy of host-pointer geometry
  size_t geom_len = 0;
  DevBuf d_direct;      // geometry DMA'd straight from page-locked caller memory
  size_t direct_len = 0;
  PinnedBuf table;      // pinned [DrawDesc x n][vbase][pbase][TargetDesc x nt]
  DevBuf d_geom, d_table, d_dstats, d_status;
  PinnedBuf h_dstats;
  HostStatus* h_status = nullptr;
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
  cudaEvent_t ev_k[RF_N_KERNELS + 1] = {};  // boundaries between the pass kernels (profiling mode)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join2 = nullptr;  // binning chain runs on the ctx's side stream
  cudaEvent_t ev_direct = nullptr;  // last asynchronous upload of page-locked caller geometry into d_direct (rf_ctx_set_geometry_async)
  bool direct_async = false;
  int profiled = 0;
  uint32_t NV = 0, NP = 0, n_tiles = 0, lt = 0;
  uint32_t n_launches = 0;
  uint32_t seq = 0;          // sequence number of the current launch of this pass
  cudaEvent_t ev_geo = nullptr;  // geometry + binning of the pass complete (geo stream)
  cudaEvent_t ev_geo_t = nullptr;  // the same instant as a timing event (RF_DEBUG_PASS timeline)
  bool first_touch = false;  // some target of the pass is cleared by k_raster / k_clear_untouched (TargetDesc::clear_flags)
  bool peer = false;         // some target of the pass replicates its colour stores into peer GPUs (rf_peer.cuh)
  bool epochs_set = false;   // barrier epochs are assigned at the first launch and reused by replays
  uint32_t epoch1 = 0, epoch2 = 0;
  uint32_t order_upper = 0;  // > 0: the pass holds a depth-sorted draw; bound on its screen triangles (rf_order.cuh)
};

}  // namespace

struct rf_target {
  rf_ctx* ctx;
  uint32_t w, h, fmt;
  bool has_depth;
  uint32_t* d_color;
  float* d_depth;
  cudaEvent_t dl_done = nullptr;  // completion of the last asynchronous download (copy stream)
  bool dl_pending = false;
  bool peer_mode = false;         // rf_target_peer_attach was called: passes drawing into it run the cross-GPU barriers
  uint32_t n_peers = 0;           // colour buffers of the same target on the other GPUs that this GPU pushes its tiles into
  uint32_t* peer_color[RF_MAX_PEERS] = {};
  bool peer_ipc[RF_MAX_PEERS] = {};  // opened with cudaIpcOpenMemHandle (closed on destroy)
  uint32_t* d_lazy = nullptr;     // lazy depth clear: RF_LAZY_WORDS words per tile (TargetDesc::lazy), zero = no tile is lazy
  bool lazy_off = false;          // the depth plane's device pointer was handed out: its memory is kept materialised from then on
};
struct rf_texture {
  rf_ctx* ctx;
  uint32_t w, h;
  uint32_t* d_data;
};
struct rf_mesh {
  rf_ctx* ctx;
  float* d_verts;
  uint32_t* d_idx;
  uint32_t n_verts, stride, n_prims, prim_kind;
};

struct rf_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_in = nullptr; // H2D of page-locked caller geometry during rf_render
  cudaStream_t copy = nullptr;   // asynchronous downloads: D2H overlaps the next pass
  cudaEvent_t ev_copy = nullptr;
  bool geometry_async = false;  // rf_ctx_set_geometry_async: page-locked caller geometry is not waited for inside rf_render
  cudaStream_t side = nullptr;   // binning chain (k_bin_alloc/scatter/sort) overlaps the span chain (k_edge_ckpt/k_walk/k_ckpt)
  bool own_stream = false;
  int sm_count = 148;
  std::string err;
  uint32_t band_y0 = 0, band_y1 = 0xFFFFFFFFu;

  PassSlot slots[kSlots];
  int cur = 0;                 // slot collecting queued draws
  std::vector<int> flight;     // slots launched and not yet validated, oldest first

  // Scratch arenas of a pass, TWO sets: consecutive passes alternate, so that the geometry stage of pass k + 1 (geo stream) can
  // run next to the rasteriser of pass k (ctx stream); a set is reused once the rasteriser that read it has finished (ev_set_done).
  struct ArenaSet {
    DevBuf cv, sv, stris, smalls, spans, tris, entries, bins, bins2, longlist, ckpts, chunks, talllist, ecks, tiles, cursors;
  } sets[2];
  cudaEvent_t ev_set_done[2] = {nullptr, nullptr};
  bool set_busy[2] = {false, false};
  uint32_t pass_seq = 0;         // sequence number of the last pass launched (device poison compares against it)
  cudaStream_t geo = nullptr;    // vertex / assembly / setup / span chain of a pass
  cudaStream_t side2 = nullptr;  // k_clear_untouched next to k_raster
  cudaEvent_t ev_pre = nullptr;
  cudaEvent_t ev_base = nullptr;  // RF_DEBUG_PASS: origin of the printed timeline
  // capacities: spans/tris/ckpts in 32-bit WORDS (record width depends on the pass's lane count),
  // entries and long spans in records
  size_t capw_stris = 0, capw_smalls = 0, capw_spans = 0, capw_tris = 0, capw_ckpts = 0, capw_ecks = 0, cap_entries = 0, cap_long = 0, cap_chunks = 0, cap_tall = 0;
  CtxStatus* d_cstatus = nullptr;
  DevBuf sdepth, ord_k32[2], ord_v[2], ord_k64[2], ord_tmp;  // Context::depth_sort (rf_order.cuh), allocated on first use
  DevBuf peer_flags;           // this GPU's barrier slots (rf_peer.cuh), written by the peers
  PeerBarrier pb{};
  bool pb_ipc[RF_MAX_PEERS + 1] = {};
  uint32_t barrier_epoch = 0;
  uint64_t replays = 0;        // passes re-launched after an arena overflow
  DevBuf bounce;               // upload/download staging on the device
  PinnedBuf h_bounce;

  // Asynchronous downloads taken since the last synchronisation: {target, host, stride, slot of the pass flushed just before}.
  // A pass replayed after arena growth re-issues the downloads that followed it, so the host buffer never keeps the frame of
  // a poisoned (no-op) pass.
  struct PendingDl { rf_target* t; void* host; size_t stride; int slot; };
  std::vector<PendingDl> pending_dl;

  rf_stats accum{};
  rf_stats last_draw{};        // stats of the last draw of the last validated pass
  uint64_t last_pass_ns = 0;
  uint32_t last_pass_launches = 0;
  int profile = 0;               // 0 off, 1 k_raster only (keeps the two-stream overlap), 2 every kernel (serialised)
  uint64_t kernel_ns[RF_N_KERNELS] = {};      // accumulated since the last query
  uint64_t kernel_launches[RF_N_KERNELS] = {};
};

namespace {

rf_status fail(rf_ctx* c, rf_status st, const char* fmt, ...) {
  if (c) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    c->err = buf;
  }
  return st;
}

#define RF_CUDA(c, expr)                                                                       \
  do {                                                                                         \
    cudaError_t e_ = (expr);                                                                   \
    if (e_ != cudaSuccess) return fail((c), RF_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)

int lt_for(uint32_t L) { return L <= 3 ? 3 : (L <= 5 ? 5 : 8); }

size_t words_cv(int lt) { return lt == 3 ? Rec<3>::CVS : lt == 5 ? Rec<5>::CVS : Rec<8>::CVS; }
size_t words_sv(int lt) { return lt == 3 ? Rec<3>::SVS : lt == 5 ? Rec<5>::SVS : Rec<8>::SVS; }
size_t words_span(int lt) { return lt == 3 ? Rec<3>::SW : lt == 5 ? Rec<5>::SW : Rec<8>::SW; }
size_t words_tri(int lt) { return lt == 3 ? Rec<3>::TW : lt == 5 ? Rec<5>::TW : Rec<8>::TW; }
size_t words_ckpt(int lt) { return lt == 3 ? Rec<3>::KW : lt == 5 ? Rec<5>::KW : Rec<8>::KW; }
size_t words_stri(int lt) { return lt == 3 ? Rec<3>::QW : lt == 5 ? Rec<5>::QW : Rec<8>::QW; }
size_t words_small(int lt) { return lt == 3 ? SmallRec<3>::W : lt == 5 ? SmallRec<5>::W : SmallRec<8>::W; }
size_t words_eck(int lt) { return lt == 3 ? Rec<3>::EW : lt == 5 ? Rec<5>::EW : Rec<8>::EW; }

struct ArenaWants {  // spans/tris/ckpts/ecks in words, the rest in records
  size_t w_spans, w_tris, w_ckpts, w_stris, entries, longs, chunks, tall, w_smalls;
};

// ---- small utility kernels --------------------------------------------------------------------
__global__ void k_fill_u32(uint32_t* p, uint32_t v, size_t n) {
  const size_t n4 = n >> 2;
  uint4* p4 = reinterpret_cast<uint4*>(p);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) p4[i] = make_uint4(v, v, v, v);
  for (size_t i = (n4 << 2) + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
// host layouts of 3- and 2-byte formats <-> uint32 containers
__global__ void k_pack_small(const uint32_t* c, uint8_t* out, size_t n, int bytes) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t v = c[i];
    for (int b = 0; b < bytes; b++) out[i * bytes + b] = (uint8_t)(v >> (8 * b));
  }
}
__global__ void k_unpack_small(const uint8_t* in, uint32_t* c, size_t n, int bytes) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t v = 0;
    for (int b = 0; b < bytes; b++) v |= (uint32_t)in[i * bytes + b] << (8 * b);
    c[i] = v;
  }
}
__global__ void k_expand_rgb(const uint8_t* in, uint32_t* out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = in[3 * i] | in[3 * i + 1] << 8 | in[3 * i + 2] << 16 | 0xFF000000u;  // Color3::to_rgba
}

int host_bytes(uint32_t fmt) { return fmt <= RF_FMT_BGRA8888 ? 4 : (fmt == RF_FMT_RGB888 ? 3 : 2); }

uint32_t pack_pixel_host(uint32_t fmt, const uint8_t c[4]) {
  const uint32_t r = c[0], g = c[1], b = c[2], a = c[3];
  switch (fmt) {
    case RF_FMT_RGBA8888: return r | g << 8 | b << 16 | a << 24;
    case RF_FMT_XRGB8888: return r << 16 | g << 8 | b;
    case RF_FMT_ARGB8888: return a | r << 8 | g << 16 | b << 24;
    case RF_FMT_BGRA8888: return b | g << 8 | r << 16 | a << 24;
    case RF_FMT_RGB888: return r | g << 8 | b << 16;
    case RF_FMT_RGB565: return ((r >> 3) & 0x1Fu) << 11 | ((g >> 2) & 0x3Fu) << 5 | ((b >> 3) & 0x1Fu);
    default: return (r >> 4) << 12 | (g >> 4) << 8 | (b >> 4) << 4 | (a >> 4);
  }
}

// ---- pass launch ----------------------------------------------------------------------------------
// Stable LSD radix sort of n (key, value) pairs on bits [0, end_bit) (rf_order.cuh): ping-pongs between the two buffer pairs,
// returns the index (0 / 1) of the pair that holds the result.
template <class K>
int rsort_pairs(rf_ctx* c, PassSlot& s, const PassParams& P, K* const k[2], uint32_t* const v[2], uint32_t n, int end_bit, cudaStream_t st) {
  const uint32_t nblk = (n + RF_RSORT_CHUNK - 1) / RF_RSORT_CHUNK;
  uint32_t* hist = static_cast<uint32_t*>(c->ord_tmp.p);
  int cur = 0;
  for (int shift = 0; shift < end_bit; shift += 8, cur ^= 1) {
    k_rsort_hist<K><<<nblk, 32, 0, st>>>(P, k[cur], n, (uint32_t)shift, hist, nblk);
    k_rsort_scan<<<1, 256, 0, st>>>(P, hist, 256u * nblk);
    k_rsort_scatter<K><<<nblk, 32, 0, st>>>(P, k[cur], k[cur ^ 1], v[cur], v[cur ^ 1], n, (uint32_t)shift, hist, nblk);
    s.n_launches += 3;
  }
  return cur;
}

// Context::depth_sort: replace the submission keys of the pass by ranks (rf_order.cuh). Runs between k_assemble and k_setup.
void launch_order(rf_ctx* c, PassSlot& s, const PassParams& P, uint32_t QW, cudaStream_t st) {
  const uint32_t n = s.order_upper;
  const unsigned g = (unsigned)std::max<size_t>(1, std::min<size_t>((n + 255) / 256, (size_t)c->sm_count * 8));
  uint32_t* k32[2] = {static_cast<uint32_t*>(c->ord_k32[0].p), static_cast<uint32_t*>(c->ord_k32[1].p)};
  uint32_t* v[2] = {static_cast<uint32_t*>(c->ord_v[0].p), static_cast<uint32_t*>(c->ord_v[1].p)};
  unsigned long long* k64[2] = {static_cast<unsigned long long*>(c->ord_k64[0].p), static_cast<unsigned long long*>(c->ord_k64[1].p)};
  k_order_init<<<g, 256, 0, st>>>(P, QW, n, k32[0], v[0]);
  const int a = rsort_pairs<uint32_t>(c, s, P, k32, v, n, 32, st);  // by submission key: restores primitive order
  uint32_t* va[2] = {v[a], v[a ^ 1]};                               // the sorted indices are the input of the second sort
  k_order_keys<<<g, 256, 0, st>>>(P, QW, n, va[0], k64[0]);
  int dbits = 1;
  while ((1ull << dbits) <= P.n_draws) dbits++;  // the padding key (all ones) stays above every draw index
  const int b = rsort_pairs<unsigned long long>(c, s, P, k64, va, n, 32 + dbits, st);  // stable by (draw, depth bits)
  k_order_apply<<<g, 256, 0, st>>>(P, QW, n, va[b]);
  s.n_launches += 3;
}

#ifndef RF_FIRST_TOUCH_CLEAR
#define RF_FIRST_TOUCH_CLEAR 1
#endif
// Blocks of the persistent rasteriser grid per SM. Below the kernel's occupancy limit the SMs keep room (registers, shared
// memory, issue slots) for the geometry kernels of the NEXT pass, which run concurrently on the geo stream.
#ifndef RF_RASTER_GRID_PER_SM
#define RF_RASTER_GRID_PER_SM 0   // 0: the occupancy limit (RasterOcc)
#endif
#ifndef RF_CLEAR_PLACE
#define RF_CLEAR_PLACE 1
#endif
#ifndef RF_FUSED_CLEAR
#define RF_FUSED_CLEAR 0   // 1: the untouched tiles of a first-touch clear are filled by k_raster's warps between their tiles (no k_clear_untouched
#endif                     // launch). Measured: k_raster grows by exactly the clear kernel's time (1.25 -> 1.50 ms, step 2.553 -> 2.546 ms; sprites and
                           // small triangles +2-3 %): the 1.5 GB of stores cost their HBM time wherever they are issued (profiles/r02_ab_fused_clear.txt)
#ifndef RF_LAZY_DEPTH
#define RF_LAZY_DEPTH 1    // a first-touch depth clear marks the untouched tiles (TargetDesc::lazy) instead of filling them
#endif
#ifndef RF_ASSEMBLE_THREADS
#define RF_ASSEMBLE_THREADS 128
#endif
#ifndef RF_SETUP_GRID_PER_SM
#define RF_SETUP_GRID_PER_SM 16
#endif
#ifndef RF_SORT_GRID_PER_SM
#define RF_SORT_GRID_PER_SM 16
#endif
template <int LT>
void launch_kernels(rf_ctx* c, PassSlot& s, const PassParams& P0) {
  PassParams P = P0;
  const bool fused_clear = RF_FUSED_CLEAR && s.first_touch && !s.peer;
  const bool clear_kernel = s.first_touch && !fused_clear;
  P.fused_clear = fused_clear ? 1u : 0u;
  const int sm = c->sm_count;
  cudaStream_t st = c->stream;
  auto blocks = [&](size_t n, int bs, int per_sm) { return (unsigned)std::max<size_t>(1, std::min<size_t>((n + bs - 1) / bs, (size_t)sm * per_sm)); };
  const int prof = c->profile;
  s.profiled = prof;
  const unsigned raster_blocks = sm * (RF_RASTER_GRID_PER_SM && LT == 3 ? RF_RASTER_GRID_PER_SM : RasterOcc<LT>::BLOCKS);
  constexpr int AT = RF_ASSEMBLE_THREADS;
  const unsigned clear_blocks = blocks((size_t)s.n_tiles * 32, 256, 8);
  if (prof == 2) {  // serialised on one stream, an event between every pair of kernels
    int ek = 0;
    auto mark = [&]() { cudaEventRecord(s.ev_k[ek++], st); };
    if (P.any_bbox) { k_objects<<<blocks(P.n_draws, 128, 8), 128, 0, st>>>(P); s.n_launches++; }
    mark(); k_vertex<LT><<<blocks(s.NV, 256, 8), 256, 0, st>>>(P);
    mark(); if (P.use_sv) k_assemble<LT, true><<<blocks(s.NP, 128, 16), 128, 0, st>>>(P); else k_assemble<LT, false><<<blocks(s.NP, 128, 16), 128, 0, st>>>(P);
    if (s.order_upper) launch_order(c, s, P, Rec<LT>::QW, st);  // counted with k_assemble
    mark(); k_setup<LT><<<sm * 8, 128, 0, st>>>(P);
    mark(); k_edge_ckpt<LT><<<sm * 4, 128, 0, st>>>(P);
    mark(); k_walk<LT><<<sm * 12, 128, 0, st>>>(P);
    mark(); k_bin_alloc<<<blocks(s.n_tiles, 256, 8), 256, 0, st>>>(P);
    mark(); k_bin_scatter<<<sm * 8, 256, 0, st>>>(P);
    mark(); k_ckpt<LT><<<sm * 8, 256, 0, st>>>(P);
    mark(); k_bin_sort_warp<<<sm * 8, RF_SORT_WARPS * 32, 0, st>>>(P);
    mark(); k_bin_sort_big<<<sm, 256, RF_SORT_BIG * 8, st>>>(P);
    mark(); if (clear_kernel) k_clear_untouched<<<clear_blocks, 256, 0, st>>>(P);
    mark();
    if (s.peer) {
      k_peer_barrier<<<1, 32, 0, st>>>(c->pb, s.epoch1);
      k_raster<LT, true><<<raster_blocks, RF_RASTER_WARPS * 32, RasterSmem<LT>::BYTES, st>>>(P);
      k_peer_barrier<<<1, 32, 0, st>>>(c->pb, s.epoch2);
      s.n_launches += 2;
    } else {
      k_raster<LT, false><<<raster_blocks, RF_RASTER_WARPS * 32, RasterSmem<LT>::BYTES, st>>>(P);
    }
    mark();
  } else {
    // Geometry of the pass on the geo stream — two dependent chains after k_setup, spans (geo) and bins (side), joined at
    // ev_geo — then the rasteriser on the ctx stream: while it runs, the geo stream already works on the next pass.
    cudaStream_t gs = c->geo, sd = c->side;
    if (P.any_bbox) { k_objects<<<blocks(P.n_draws, 128, 8), 128, 0, gs>>>(P); s.n_launches++; }
    k_vertex<LT><<<blocks(s.NV, 256, 8), 256, 0, gs>>>(P);
    if (P.use_sv) k_assemble<LT, true><<<blocks(s.NP, AT, 16 * 128 / AT), AT, 0, gs>>>(P); else k_assemble<LT, false><<<blocks(s.NP, AT, 16 * 128 / AT), AT, 0, gs>>>(P);
    if (s.order_upper) launch_order(c, s, P, Rec<LT>::QW, gs);
    k_setup<LT><<<sm * RF_SETUP_GRID_PER_SM, 128, 0, gs>>>(P);
    cudaEventRecord(s.ev_fork, gs);
    cudaStreamWaitEvent(sd, s.ev_fork, 0);
    // The untouched tiles of first-touch-cleared targets are known once k_setup has counted the bins. RF_CLEAR_PLACE 1: they
    // are filled on the ctx stream while the bins are sorted; 2: after k_raster, i.e. next to the geometry stage of the NEXT
    // pass (latency-bound kernels on the geo stream) — never next to k_raster itself, which that slowed by 14 %
    // (profiles/r02_ab_clear_placement.txt). With peers the clear must precede the cross-GPU barrier: always placement 1.
    if (clear_kernel && (RF_CLEAR_PLACE == 1 || s.peer)) {
      cudaStreamWaitEvent(st, s.ev_fork, 0);
      k_clear_untouched<<<clear_blocks, 256, 0, st>>>(P);
    }
    k_bin_alloc<<<blocks(s.n_tiles, 256, 8), 256, 0, sd>>>(P);
    k_bin_scatter<<<sm * 4, 256, 0, sd>>>(P);
    k_bin_sort_warp<<<sm * RF_SORT_GRID_PER_SM, RF_SORT_WARPS * 32, 0, sd>>>(P);
    k_bin_sort_big<<<sm, 256, RF_SORT_BIG * 8, sd>>>(P);
    cudaEventRecord(s.ev_join, sd);
    k_edge_ckpt<LT><<<sm * 4, 128, 0, gs>>>(P);
    k_walk<LT><<<sm * 12, 128, 0, gs>>>(P);
    k_ckpt<LT><<<sm * 8, 256, 0, gs>>>(P);
    cudaStreamWaitEvent(gs, s.ev_join, 0);
    cudaEventRecord(s.ev_geo, gs);
    if (s.ev_geo_t) cudaEventRecord(s.ev_geo_t, gs);
    cudaStreamWaitEvent(st, s.ev_geo, 0);
    if (s.peer) { k_peer_barrier<<<1, 32, 0, st>>>(c->pb, s.epoch1); s.n_launches++; }  // every peer has cleared its copy of the frame
    if (prof == 1) cudaEventRecord(s.ev_k[RF_N_KERNELS - 1], st);
    if (s.peer) k_raster<LT, true><<<raster_blocks, RF_RASTER_WARPS * 32, RasterSmem<LT>::BYTES, st>>>(P);
    else k_raster<LT, false><<<raster_blocks, RF_RASTER_WARPS * 32, RasterSmem<LT>::BYTES, st>>>(P);
    if (prof == 1) cudaEventRecord(s.ev_k[RF_N_KERNELS], st);
    if (clear_kernel && RF_CLEAR_PLACE == 2 && !s.peer) k_clear_untouched<<<clear_blocks, 256, 0, st>>>(P);
    if (s.peer) { k_peer_barrier<<<1, 32, 0, st>>>(c->pb, s.epoch2); s.n_launches++; }  // every peer's stores into this GPU have landed
  }
  s.n_launches += RF_N_KERNELS - (clear_kernel ? 0 : 1);
}

// Any growth frees memory that in-flight kernels might still use -> callers guarantee idleness.
rf_status ensure_arenas(rf_ctx* c, int lt, size_t nv, size_t n_tiles, const ArenaWants& w) {
  // every buffer at the larger of its present capacity and the wanted one (reserve() is a no-op for a buffer that is large enough)
  const size_t stris = std::max(c->capw_stris, w.w_stris), smalls = std::max(c->capw_smalls, w.w_smalls), spans = std::max(c->capw_spans, w.w_spans),
               tris = std::max(c->capw_tris, w.w_tris), ckpts = std::max(c->capw_ckpts, w.w_ckpts), entries = std::max(c->cap_entries, w.entries),
               longs = std::max(c->cap_long, w.longs), chunks = std::max(c->cap_chunks, w.chunks), tall = std::max(c->cap_tall, w.tall);
  auto reserve_set = [&](rf_ctx::ArenaSet& a) {
    return a.cv.reserve(nv * words_cv(lt) * 4 + 64) && a.sv.reserve(nv * words_sv(lt) * 4 + 64) && a.tiles.reserve(n_tiles * kTileArrays * 4 + 64) &&
           a.cursors.reserve(64) && a.stris.reserve(stris * 4) && a.smalls.reserve(smalls * 4) && a.spans.reserve(spans * 4) && a.tris.reserve(tris * 4) &&
           a.ckpts.reserve(ckpts * 4) && a.entries.reserve(entries * 16) && a.bins.reserve(entries * 8) && a.bins2.reserve(entries * 8) &&
           a.longlist.reserve(longs * 8) && a.chunks.reserve(chunks * 16) && a.ecks.reserve(chunks * Rec<8>::EW * 4) && a.talllist.reserve(tall * 4);
  };
  bool ok = reserve_set(c->sets[0]) && reserve_set(c->sets[1]);
  if (!ok) {
    // A buffer is freed before its larger replacement is allocated, so the peak is the new footprint; a failure here was seen
    // once on a 2-GPU box right after other CUDA processes had exited. Let the device settle, give everything back, start over.
    cudaGetLastError();
    cudaDeviceSynchronize();
    for (auto& a : c->sets) {
      a.cv.release(); a.sv.release(); a.stris.release(); a.smalls.release(); a.spans.release(); a.tris.release(); a.entries.release(); a.bins.release();
      a.bins2.release(); a.longlist.release(); a.ckpts.release(); a.chunks.release(); a.talllist.release(); a.ecks.release(); a.tiles.release(); a.cursors.release();
    }
    ok = reserve_set(c->sets[0]) && reserve_set(c->sets[1]);
  }
  if (!ok) {
    size_t fr = 0, tot = 0;
    const cudaError_t le = cudaGetLastError();
    cudaMemGetInfo(&fr, &tot);
    c->capw_stris = c->capw_smalls = c->capw_spans = c->capw_tris = c->capw_ckpts = c->cap_entries = c->cap_long = c->cap_chunks = c->cap_tall = 0;  // nothing is certain now
    return fail(c, RF_E_NOMEM, "pass arenas (%s; free %zu of %zu MiB; want spans %zu tris %zu ckpts %zu stris %zu smalls %zu MiB, entries %zu, verts %zu)", cudaGetErrorString(le),
                fr >> 20, tot >> 20, spans >> 18, tris >> 18, ckpts >> 18, stris >> 18, smalls >> 18, entries, nv);
  }
  c->capw_stris = stris; c->capw_smalls = smalls; c->capw_spans = spans; c->capw_tris = tris; c->capw_ckpts = ckpts;
  c->cap_entries = entries; c->cap_long = longs; c->cap_chunks = chunks; c->cap_tall = tall;
  return RF_OK;
}

bool arenas_cover(const rf_ctx* c, const ArenaWants& w) {
  return w.w_stris <= c->capw_stris && w.w_smalls <= c->capw_smalls && w.w_spans <= c->capw_spans && w.w_tris <= c->capw_tris && w.w_ckpts <= c->capw_ckpts &&
         w.entries <= c->cap_entries && w.longs <= c->cap_long && w.chunks <= c->cap_chunks && w.tall <= c->cap_tall;
}

rf_status wait_idle(rf_ctx* c) {
  RF_CUDA(c, cudaStreamSynchronize(c->geo));
  RF_CUDA(c, cudaStreamSynchronize(c->side));
  RF_CUDA(c, cudaStreamSynchronize(c->stream));
  RF_CUDA(c, cudaStreamSynchronize(c->side2));
  c->set_busy[0] = c->set_busy[1] = false;
  return RF_OK;
}

// Launch (or re-launch) the pass stored in slot `si`.
// ---- lazy depth clear (TargetDesc::lazy)
uint32_t lazy_tiles(const rf_target* t) { return ((t->w + RF_TILE - 1) >> RF_TILE_SHIFT) * ((t->h + RF_TILE - 1) >> RF_TILE_SHIFT); }
uint32_t lazy_words(const rf_target* t) { return lazy_tiles(t) * RF_LAZY_WORDS; }
// writes the marked tiles of the target out, on the ctx stream (ordered after every pass launched so far)
void lazy_materialize(rf_ctx* c, rf_target* t) {
  if (!t->d_lazy || t->lazy_off) return;
  const uint32_t nt = lazy_tiles(t);
  const unsigned grid = std::max(1u, std::min((nt + 7u) / 8u, (uint32_t)c->sm_count * 8u));
  const uint32_t tiles_x = (t->w + RF_TILE - 1) >> RF_TILE_SHIFT;
  k_lazy_materialize<<<grid, 256, 0, c->stream>>>(t->d_depth, t->d_lazy, t->w, t->h, tiles_x, nt);
}

rf_status launch_pass(rf_ctx* c, int si) {
  PassSlot& s = c->slots[si];
  const size_t nd = s.draws.size();
  // ---- targets and tiles
  s.targets.clear();
  for (auto& q : s.draws) {
    auto it = std::find(s.targets.begin(), s.targets.end(), q.target);
    if (it == s.targets.end()) { s.targets.push_back(q.target); q.desc.target = (uint32_t)s.targets.size() - 1; }
    else q.desc.target = (uint32_t)(it - s.targets.begin());
  }
  const size_t nt = s.targets.size();
  uint32_t maxL = 0;
  for (auto& q : s.draws) maxL = std::max(maxL, q.desc.L);
  const int lt = lt_for(maxL);
  s.lt = lt;

  // clears of targets the pass also draws into are first-touch clears (see the target table below): no ClearDesc
  auto drawn = [&](const rf_target* t) { return RF_FIRST_TOUCH_CLEAR && std::find(s.targets.begin(), s.targets.end(), t) != s.targets.end(); };
  size_t ncl = 0;
  bool first_touch = false;
  for (const QueuedClear& qc : s.clears) {
    const bool ft = drawn(qc.target);
    if (qc.has_color && !(ft && !qc.target->peer_mode)) ncl++;
    if (qc.has_depth && !ft) ncl++;
    first_touch = first_touch || (ft && (qc.has_depth || (qc.has_color && !qc.target->peer_mode)));
  }
  s.first_touch = first_touch;
  const size_t table_bytes = nd * sizeof(DrawDesc) + 2 * (nd + 1) * 4 + nt * sizeof(TargetDesc) + ncl * sizeof(ClearDesc) + 64;
  if (!s.table.reserve(table_bytes)) return fail(c, RF_E_NOMEM, "pinned table");
  // idle device needed before any reallocation of buffers a previous launch of this slot used
  bool need_idle = s.d_table.cap < table_bytes || s.d_geom.cap < s.geom_len || s.d_dstats.cap < nd * sizeof(DrawStats);
  uint32_t nv = 0, np = 0, ntiles = 0;
  for (auto& q : s.draws) { nv += q.desc.n_verts; np += q.desc.n_prims; }
  for (auto* t : s.targets) ntiles += ((t->w + RF_TILE - 1) / RF_TILE) * ((t->h + RF_TILE - 1) / RF_TILE);
  s.NV = nv; s.NP = np; s.n_tiles = ntiles;
  // initial arena sizes (grown on demand by validate_all after an overflowing pass)
  ArenaWants want{std::max<size_t>(c->capw_spans, (size_t)8 << 20), std::max<size_t>(c->capw_tris, (size_t)8 << 20),
                  std::max<size_t>(c->capw_ckpts, (size_t)2 << 20), std::max<size_t>(c->capw_stris, (size_t)8 << 20),
                  std::max<size_t>(c->cap_entries, (size_t)1 << 20), std::max<size_t>(c->cap_long, (size_t)1 << 19),
                  std::max<size_t>(c->cap_chunks, (size_t)1 << 20), std::max<size_t>(c->cap_tall, (size_t)1 << 18),
                  std::max<size_t>(c->capw_smalls, (size_t)8 << 20)};
  for (auto& a : c->sets)
    need_idle = need_idle || a.cv.cap < (size_t)nv * words_cv(lt) * 4 + 64 || a.sv.cap < (size_t)nv * words_sv(lt) * 4 + 64 ||
                a.tiles.cap < (size_t)ntiles * kTileArrays * 4 + 64 || a.cursors.cap < 64;
  need_idle = need_idle || !arenas_cover(c, want);
  // Context::depth_sort: bound on the pass's screen triangles (a clipped triangle fans into <= 7) and the sort buffers
  uint64_t order_bound = 0;
  bool any_sort = false;
  for (auto& q : s.draws) {
    order_bound += (uint64_t)q.desc.n_prims * (q.desc.prim_kind == RF_PRIM_EDGES ? 1 : 7);
    any_sort = any_sort || ((q.desc.flags >> RF_F_DSORT_SHIFT) & RF_F_DSORT_MASK) != 0;
  }
  const size_t cap_stris_el = std::min<size_t>(std::max(want.w_stris, c->capw_stris) / words_stri(lt), 0x1FFFFFF0u);
  const uint32_t order_upper = any_sort ? (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(order_bound, cap_stris_el)) : 0u;
  size_t order_tmp = 0;
  if (order_upper) {
    order_tmp = (size_t)((order_upper + RF_RSORT_CHUNK - 1) / RF_RSORT_CHUNK) * 256 * 4 + 256;  // digit histograms of the radix sort
    need_idle = need_idle || c->sdepth.cap < cap_stris_el * 4 || c->ord_k32[0].cap < (size_t)order_upper * 4 || c->ord_tmp.cap < order_tmp;
  }
  s.order_upper = order_upper;
  if (need_idle) { rf_status st = wait_idle(c); if (st) return st; }
  if (order_upper) {
    bool ok = c->sdepth.reserve(cap_stris_el * 4) && c->ord_tmp.reserve(order_tmp);
    for (int k = 0; k < 2; k++) ok = ok && c->ord_k32[k].reserve((size_t)order_upper * 4) && c->ord_v[k].reserve((size_t)order_upper * 4) && c->ord_k64[k].reserve((size_t)order_upper * 8);
    if (!ok) return fail(c, RF_E_NOMEM, "depth-sort buffers");
  }
  if (!s.d_table.reserve(table_bytes) || !s.d_geom.reserve(std::max<size_t>(s.geom_len, 16)) ||
      !s.d_dstats.reserve(std::max<size_t>(nd * sizeof(DrawStats), 16)) || !s.d_status.reserve(sizeof(PassStatus)) ||
      !s.h_dstats.reserve(std::max<size_t>(nd * sizeof(DrawStats), 16)))
    return fail(c, RF_E_NOMEM, "pass buffers");
  { rf_status st = ensure_arenas(c, lt, nv, ntiles, want); if (st) return st; }

  // ---- build the table
  uint8_t* tb = s.table.p;
  DrawDesc* h_draws = reinterpret_cast<DrawDesc*>(tb);
  uint32_t* h_vbase = reinterpret_cast<uint32_t*>(tb + nd * sizeof(DrawDesc));
  uint32_t* h_pbase = h_vbase + (nd + 1);
  size_t toff = nd * sizeof(DrawDesc) + 2 * (nd + 1) * 4;
  toff = (toff + 15) & ~size_t(15);
  TargetDesc* h_targets = reinterpret_cast<TargetDesc*>(tb + toff);
  uint32_t vb = 0, pb = 0;
  bool uniform_verts = true, uniform_prims = true;  // frame batch: every draw has the same vertex / primitive count
  for (size_t i = 0; i < nd; i++) {
    QueuedDraw& q = s.draws[i];
    DrawDesc d = q.desc;
    if (q.verts_off != SIZE_MAX) {
      uint8_t* gbase = static_cast<uint8_t*>(q.direct ? s.d_direct.p : s.d_geom.p);
      d.verts = reinterpret_cast<const float*>(gbase + q.verts_off);
      d.indices = reinterpret_cast<const uint32_t*>(gbase + q.idx_off);
    }
    h_draws[i] = d;
    h_vbase[i] = vb; h_pbase[i] = pb;
    vb += d.n_verts; pb += d.n_prims;
    if (d.n_verts != h_draws[0].n_verts) uniform_verts = false;
    if (d.n_prims != h_draws[0].n_prims) uniform_prims = false;
  }
  h_vbase[nd] = vb; h_pbase[nd] = pb;
  uint32_t tile_base = 0, tiles_per_target = 0;
  bool uniform_tiles = nt > 0;
  for (size_t i = 0; i < nt; i++) {
    rf_target* t = s.targets[i];
    TargetDesc& T = h_targets[i];
    T.color = t->d_color; T.depth = t->has_depth ? t->d_depth : nullptr;
    T.w = t->w; T.h = t->h; T.fmt = t->fmt;
    T.tiles_x = (t->w + RF_TILE - 1) / RF_TILE; T.tiles_y = (t->h + RF_TILE - 1) / RF_TILE;
    T.tile_base = tile_base; tile_base += T.tiles_x * T.tiles_y;
    if (i == 0) tiles_per_target = T.tiles_x * T.tiles_y;
    else if (T.tiles_x * T.tiles_y != tiles_per_target) uniform_tiles = false;
    T.band_y0 = std::min(c->band_y0, t->h); T.band_y1 = std::min(c->band_y1, t->h);
    T.n_peers = t->n_peers;
    T.clear_flags = 0; T.clear_color = 0; T.clear_zbits = 0;
    T.lazy = (t->has_depth && t->d_lazy && !t->lazy_off) ? t->d_lazy : nullptr;
    // new lazy marks: not under sort-first sharding (a band owns part of a tile's rows; peers gather depth planes as memory)
    const bool lazy_ok = T.lazy != nullptr && !t->peer_mode && c->band_y0 == 0 && c->band_y1 >= t->h;
    // First-touch clear: a Frame::clear recorded at the head of this pass for a target the pass draws into is carried out by
    // k_raster (touched tiles) and k_clear_untouched (the others). With peers attached the other GPUs store their bands into
    // this colour buffer, so its colour plane is still cleared as a whole, before the first cross-GPU barrier.
    for (const QueuedClear& qc : s.clears) {
      if (qc.target != t || !RF_FIRST_TOUCH_CLEAR) continue;
      if (qc.has_color && !t->peer_mode) { T.clear_flags |= RF_CLEAR_COLOR; T.clear_color = qc.color; }
      if (qc.has_depth) { T.clear_flags |= RF_CLEAR_DEPTH | (lazy_ok ? RF_CLEAR_LAZY : 0u); T.clear_zbits = qc.zbits; }
    }
    for (uint32_t p = 0; p < RF_MAX_PEERS; p++) T.peer_color[p] = p < t->n_peers ? t->peer_color[p] : nullptr;
    if (t->peer_mode && nd) s.peer = true;
  }
  if (s.peer && !s.epochs_set) { s.epoch1 = ++c->barrier_epoch; s.epoch2 = ++c->barrier_epoch; s.epochs_set = true; }

  const size_t coff = (toff + nt * sizeof(TargetDesc) + 15) & ~size_t(15);
  ClearDesc* h_clears = reinterpret_cast<ClearDesc*>(tb + coff);
  {
    size_t k = 0;
    for (const QueuedClear& qc : s.clears) {
      // under sort-first sharding only the rows of this GPU's band are cleared (the others are never rasterised here)
      const uint32_t y0 = std::min(c->band_y0, qc.target->h), y1 = std::min(c->band_y1, qc.target->h);
      const size_t first = (size_t)y0 * qc.target->w;
      const unsigned long long n = (unsigned long long)(y1 > y0 ? y1 - y0 : 0) * qc.target->w;
      // with peers attached the other GPUs store THEIR bands into this buffer: the colour clear covers every row
      const bool ft = drawn(qc.target);
      if (qc.has_color && qc.target->peer_mode) h_clears[k++] = ClearDesc{qc.target->d_color, (unsigned long long)qc.target->w * qc.target->h, qc.color, 0u, nullptr};
      else if (qc.has_color && !ft) h_clears[k++] = ClearDesc{qc.target->d_color + first, n, qc.color, 0u, nullptr};
      if (qc.has_depth && !ft) {
        // the plane is written as memory: no tile of it stays lazy (a band clears only its rows: the marked tiles are written out first)
        rf_target* qt = qc.target;
        const bool lazy = qt->d_lazy && !qt->lazy_off, whole = y0 == 0 && y1 >= qt->h;
        if (lazy && !whole) lazy_materialize(c, qt);
        h_clears[k++] = ClearDesc{reinterpret_cast<uint32_t*>(qt->d_depth) + first, n, qc.zbits, lazy && whole ? lazy_words(qt) : 0u, qt->d_lazy};
      }
    }
  }

  cudaStream_t st = c->stream;
  // Streams. Everything that touches a TARGET (clears, k_raster, k_clear_untouched) is ordered on the ctx stream; the geometry
  // of the pass — table / geometry uploads included — runs on the geo stream and only has to wait until the rasteriser that
  // last read this arena set is done. In the serialised profiling mode everything is on the ctx stream.
  s.seq = ++c->pass_seq;
  const int set = (int)(s.seq & 1u);
  rf_ctx::ArenaSet& A = c->sets[set];
  cudaStream_t gs = c->profile == 2 ? st : c->geo;
  // a target that is still being downloaded on the copy stream must not be overwritten yet
  for (rf_target* t : s.targets) if (t->dl_pending) { RF_CUDA(c, cudaStreamWaitEvent(st, t->dl_done, 0)); t->dl_pending = false; }
  for (const QueuedClear& qc : s.clears) if (qc.target->dl_pending) { RF_CUDA(c, cudaStreamWaitEvent(st, qc.target->dl_done, 0)); qc.target->dl_pending = false; }
  if (c->set_busy[set] && gs != st) RF_CUDA(c, cudaStreamWaitEvent(gs, c->ev_set_done[set], 0));
  if (getenv("RF_DEBUG_PASS")) {
    if (!c->ev_base) { cudaEventCreate(&c->ev_base); cudaEventRecord(c->ev_base, st); }
    if (!s.ev_geo_t) cudaEventCreate(&s.ev_geo_t);
  }
  RF_CUDA(c, cudaEventRecord(s.ev_start, gs));
  s.n_launches = 0;
  RF_CUDA(c, cudaMemcpyAsync(s.d_table.p, tb, coff + ncl * sizeof(ClearDesc), cudaMemcpyHostToDevice, gs));
  if (ncl || nd == 0) {
    if (gs != st) { RF_CUDA(c, cudaEventRecord(s.ev_fork, gs)); RF_CUDA(c, cudaStreamWaitEvent(st, s.ev_fork, 0)); }  // the table
    if (ncl) {
      const unsigned gx = (unsigned)std::max<size_t>(1, std::min<size_t>((size_t)c->sm_count * 8 / ncl + 1, 1024));
      k_clear_multi<<<dim3(gx, (unsigned)ncl), 256, 0, st>>>(reinterpret_cast<const ClearDesc*>(static_cast<uint8_t*>(s.d_table.p) + coff), c->d_cstatus, s.seq);
      s.n_launches++;
    }
  }
  if (nd == 0) {
    RF_CUDA(c, cudaMemsetAsync(s.d_status.p, 0, sizeof(PassStatus), st));
    RF_CUDA(c, cudaMemcpyAsync(&s.h_status->status, s.d_status.p, sizeof(PassStatus), cudaMemcpyDeviceToHost, st));
    RF_CUDA(c, cudaEventRecord(s.ev_stop, st));
    RF_CUDA(c, cudaEventRecord(c->ev_set_done[set], st));
    c->set_busy[set] = true;
    s.in_flight = true;
    return RF_OK;
  }
  if (s.direct_async) RF_CUDA(c, cudaStreamWaitEvent(gs, s.ev_direct, 0));  // asynchronous uploads of page-locked geometry
  if (s.geom_len) RF_CUDA(c, cudaMemcpyAsync(s.d_geom.p, s.geom.p, s.geom_len, cudaMemcpyHostToDevice, gs));
  RF_CUDA(c, cudaMemsetAsync(s.d_status.p, 0, sizeof(PassStatus), gs));
  RF_CUDA(c, cudaMemsetAsync(s.d_dstats.p, 0, std::max<size_t>(nd * sizeof(DrawStats), 16), gs));
  RF_CUDA(c, cudaMemsetAsync(A.tiles.p, 0, (size_t)ntiles * 20, gs));
  RF_CUDA(c, cudaMemsetAsync(A.cursors.p, 0, 64, gs));

  PassParams P{};
  uint8_t* dt = static_cast<uint8_t*>(s.d_table.p);
  P.draws = reinterpret_cast<const DrawDesc*>(dt);
  P.vbase = reinterpret_cast<const uint32_t*>(dt + nd * sizeof(DrawDesc));
  P.pbase = P.vbase + (nd + 1);
  P.targets = reinterpret_cast<const TargetDesc*>(dt + toff);
  P.seq = s.seq;
  P.n_draws = (uint32_t)nd; P.n_targets = (uint32_t)nt; P.NV = nv; P.NP = np; P.n_tiles = ntiles;
  P.verts_per_draw = nd && uniform_verts ? h_draws[0].n_verts : 0u;  // 0 also when the draws are empty: binary search
  P.prims_per_draw = nd && uniform_prims ? h_draws[0].n_prims : 0u;
  P.tiles_per_target = uniform_tiles ? tiles_per_target : 0u;  // frame batches: k_raster finds a tile's target by one division
  for (auto& q : s.draws) if (q.desc.flags & RF_F_BBOX) P.any_bbox = 1;
  P.use_sv = 1;
  for (auto& q : s.draws) if (!(q.desc.flags & RF_F_SV)) P.use_sv = 0;
  P.cv = static_cast<float*>(A.cv.p);
  P.sv = static_cast<float*>(A.sv.p);
  P.stris = static_cast<uint32_t*>(A.stris.p);
  P.cap_stris = (uint32_t)std::min<size_t>(c->capw_stris / words_stri(lt), 0x1FFFFFF0u);
  P.smalls = static_cast<uint32_t*>(A.smalls.p);
  P.cap_smalls = (uint32_t)std::min<size_t>(c->capw_smalls / words_small(lt), 0x7FFFFFF0u);
  P.sdepth = s.order_upper ? static_cast<uint32_t*>(c->sdepth.p) : nullptr;
  P.spans = static_cast<uint32_t*>(A.spans.p);
  P.tris = static_cast<uint32_t*>(A.tris.p);
  P.entries = static_cast<uint4*>(A.entries.p);
  P.bins = static_cast<unsigned long long*>(A.bins.p);
  P.bins2 = static_cast<unsigned long long*>(A.bins2.p);
  P.longlist = static_cast<uint2*>(A.longlist.p);
  P.ckpts = static_cast<uint32_t*>(A.ckpts.p);
  // capacities in records of THIS pass's width
  P.cap_spans = (uint32_t)std::min<size_t>(c->capw_spans / words_span(lt), 0xFFFFFFF0u);
  P.cap_tris = (uint32_t)std::min<size_t>(c->capw_tris / words_tri(lt), 0x7FFFFFF0u);
  P.cap_ckpts = (uint32_t)std::min<size_t>(c->capw_ckpts / words_ckpt(lt), 0xFFFFFFF0u);
  P.cap_entries = (uint32_t)std::min<size_t>(c->cap_entries, 0xFFFFFFF0u);
  P.cap_long = (uint32_t)std::min<size_t>(c->cap_long, 0xFFFFFFF0u);
  P.chunks = static_cast<uint4*>(A.chunks.p);
  P.talllist = static_cast<uint32_t*>(A.talllist.p);
  P.ecks = static_cast<uint32_t*>(A.ecks.p);
  P.cap_chunks = (uint32_t)std::min<size_t>(c->cap_chunks, 0xFFFFFFF0u);
  P.cap_tall = (uint32_t)std::min<size_t>(c->cap_tall, 0xFFFFFFF0u);
  P.cap_ecks = P.cap_chunks;
  uint32_t* ta = static_cast<uint32_t*>(A.tiles.p);
  P.tile_cnt = ta; P.tile_off = ta + ntiles; P.tile_fill = ta + 2 * (size_t)ntiles;
  P.worklist = ta + 3 * (size_t)ntiles; P.worklist_big = ta + 4 * (size_t)ntiles; P.worklist_heavy = ta + 5 * (size_t)ntiles;
  P.cursors = static_cast<uint32_t*>(A.cursors.p);
  P.dstats = static_cast<DrawStats*>(s.d_dstats.p);
  P.status = static_cast<PassStatus*>(s.d_status.p);
  P.cstatus = c->d_cstatus;

  if (lt == 3) launch_kernels<3>(c, s, P);
  else if (lt == 5) launch_kernels<5>(c, s, P);
  else launch_kernels<8>(c, s, P);
  RF_CUDA(c, cudaGetLastError());

  RF_CUDA(c, cudaMemcpyAsync(&s.h_status->status, s.d_status.p, sizeof(PassStatus), cudaMemcpyDeviceToHost, st));
  if (nd) RF_CUDA(c, cudaMemcpyAsync(s.h_dstats.p, s.d_dstats.p, nd * sizeof(DrawStats), cudaMemcpyDeviceToHost, st));
  RF_CUDA(c, cudaEventRecord(s.ev_stop, st));
  RF_CUDA(c, cudaEventRecord(c->ev_set_done[set], st));
  c->set_busy[set] = true;
  s.in_flight = true;
  return RF_OK;
}

void reset_slot(PassSlot& s) {
  s.clears.clear();
  s.draws.clear();
  s.prims_queued = 0;
  s.geom_len = 0;
  s.direct_len = 0;
  s.in_flight = false;
  s.peer = false; s.epochs_set = false;
  s.direct_async = false;
}

// The D2H copy of rf_target_download_color_async: on the copy stream, after everything queued so far on the ctx stream.
rf_status issue_download(rf_ctx* c, rf_target* t, void* host, size_t stride) {
  RF_CUDA(c, cudaEventRecord(c->ev_copy, c->stream));
  RF_CUDA(c, cudaStreamWaitEvent(c->copy, c->ev_copy, 0));
  RF_CUDA(c, cudaMemcpy2DAsync(host, stride * 4, t->d_color, (size_t)t->w * 4, (size_t)t->w * 4, t->h, cudaMemcpyDeviceToHost, c->copy));
  if (!t->dl_done) RF_CUDA(c, cudaEventCreateWithFlags(&t->dl_done, cudaEventDisableTiming));
  RF_CUDA(c, cudaEventRecord(t->dl_done, c->copy));
  t->dl_pending = true;
  return RF_OK;
}

// Wait for every pass in flight, replay overflowed ones with larger arenas, fold Stats.
// `only_oldest`: stop once the oldest pass has been validated (the others stay in flight: the pipeline is not drained).
rf_status validate_all(rf_ctx* c, bool only_oldest = false) {
  rf_status result = RF_OK;
  const size_t stop_at = only_oldest && !c->flight.empty() ? c->flight.size() - 1 : 0;
  while (c->flight.size() > stop_at) {
    const int si = c->flight.front();
    PassSlot& s = c->slots[si];
    RF_CUDA(c, cudaEventSynchronize(s.ev_stop));
    const PassStatus ps = s.h_status->status;
    if (getenv("RF_DEBUG_PASS"))
      fprintf(stderr, "[rf pass] draws %zu stris %llu small %llu tris %llu spans %llu entries %llu chunks %llu long %llu ckpts %llu work %u max_bin %u overflow %u error %u\n",
              s.draws.size(), ps.stris_needed.v, ps.small_needed.v, ps.tris_needed.v, ps.spans_needed.v, ps.entries_needed.v, ps.chunks_needed.v,
              ps.long_needed.v, ps.ckpts_needed.v, ps.n_work, ps.max_bin, ps.overflow, ps.error);
    if (getenv("RF_DEBUG_PASS") && c->ev_base && s.ev_geo_t && !s.draws.empty()) {
      float a = 0, b = 0, r0 = 0, r1 = 0, e = 0;
      cudaEventElapsedTime(&a, c->ev_base, s.ev_start); cudaEventElapsedTime(&b, c->ev_base, s.ev_geo_t); cudaEventElapsedTime(&e, c->ev_base, s.ev_stop);
      if (s.profiled == 1) { cudaEventElapsedTime(&r0, c->ev_base, s.ev_k[RF_N_KERNELS - 1]); cudaEventElapsedTime(&r1, c->ev_base, s.ev_k[RF_N_KERNELS]); }
      fprintf(stderr, "[rf timeline] seq %u: geometry %.3f .. %.3f ms, raster %.3f .. %.3f, done %.3f\n", s.seq, a, b, r0, r1, e);
      cudaGetLastError();  // an event of this debug printout that was never recorded must not surface as the next call's error
    }
    if (ps.overflow) {
      // Every later pass in flight was a no-op (device poison). Grow and replay from here, in order.
      { rf_status st = wait_idle(c); if (st) return st; }
      const int lt = (int)s.lt;
      auto grow = [](size_t cap, unsigned long long need) { return need > cap ? (size_t)(need + need / 4 + 4096) : cap; };
      // counters downstream of an overflowed stage are incomplete: guess them from the span count
      const bool early = ps.stris_needed * words_stri(lt) > c->capw_stris || ps.small_needed * words_small(lt) > c->capw_smalls || ps.spans_needed * words_span(lt) > c->capw_spans || ps.tris_needed * words_tri(lt) > c->capw_tris ||
                         ps.chunks_needed > c->cap_chunks || ps.entries_needed > c->cap_entries;
      ArenaWants w{grow(c->capw_spans, ps.spans_needed * words_span(lt)), grow(c->capw_tris, ps.tris_needed * words_tri(lt)),
                   grow(c->capw_ckpts, std::max<unsigned long long>(ps.ckpts_needed, early ? ps.spans_needed / 8 : 0) * words_ckpt(lt)),
                   grow(c->capw_stris, ps.stris_needed * words_stri(lt)), grow(c->cap_entries, ps.entries_needed),
                   grow(c->cap_long, std::max<unsigned long long>(ps.long_needed, early ? ps.spans_needed / 8 : 0)),
                   grow(c->cap_chunks, ps.chunks_needed), grow(c->cap_tall, ps.tall_needed), grow(c->capw_smalls, ps.small_needed * words_small(lt))};
      { rf_status st = ensure_arenas(c, lt, s.NV, s.n_tiles, w); if (st) return st; }
      RF_CUDA(c, cudaMemsetAsync(c->d_cstatus, 0xFF, sizeof(CtxStatus), c->stream));
      RF_CUDA(c, cudaStreamSynchronize(c->stream));
      std::vector<int> replay = c->flight;
      c->replays += replay.size();
      for (int r : replay) {
        rf_status st = launch_pass(c, r);
        if (st) return st;
        for (const rf_ctx::PendingDl& d : c->pending_dl) if (d.slot == r) { st = issue_download(c, d.t, d.host, d.stride); if (st) return st; }
      }
      continue;  // re-validate the same front slot
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, s.ev_start, s.ev_stop);
    if (s.profiled && !s.draws.empty()) {
      for (int k = (s.profiled == 2 ? 0 : RF_N_KERNELS - 1); k < RF_N_KERNELS; k++) {
        float kms = 0.f;
        if (cudaEventElapsedTime(&kms, s.ev_k[k], s.ev_k[k + 1]) == cudaSuccess) { c->kernel_ns[k] += (uint64_t)(kms * 1e6); c->kernel_launches[k] += 1; }
      }
    }
    const uint64_t ns = (uint64_t)(ms * 1e6);
    c->last_pass_ns = ns;
    c->last_pass_launches = s.n_launches;
    if (ps.error) {
      if (ps.error & RF_ERRBIT_INDEX_OOB) result = fail(c, RF_E_INDEX_OOB, "vertex index out of bounds (render/prim.rs:17-19 panics)");
      else if (ps.error & RF_ERRBIT_TARGET_OOB) result = fail(c, RF_E_TARGET_OOB, "scanline outside the render target (render/target.rs:148,173 panics)");
      else if (ps.error & RF_ERRBIT_TEXEL_OOB) result = fail(c, RF_E_BAD_TEXTURE, "SamplerOnce: texture coordinate outside the texture (render/tex.rs:343-356 panics)");
      else if (ps.error & RF_ERRBIT_INTERNAL) result = fail(c, RF_E_CUDA, "internal: span outside its triangle's tile bounding box");
      else result = fail(c, RF_E_NOMEM, "a 32x32 tile is overlapped by more than %u triangles of one pass (max_bin=%u)", RF_SORT_BIG, ps.max_bin);
    } else {
      const DrawStats* ds = reinterpret_cast<const DrawStats*>(s.h_dstats.p);
      for (size_t i = 0; i < s.draws.size(); i++) {
        if (s.draws[i].desc.flags & RF_F_BBOX) {  // crates.rs:101,118-131: objs.i every object, objs.o the ones rendered
          c->accum.objs_i += 1;
          if (ds[i].hidden) { if (i + 1 == s.draws.size()) { c->last_draw = rf_stats{}; c->last_draw.objs_i = 1; } continue; }
          c->accum.objs_o += 1;
        }
        rf_stats d{};
        if (s.draws[i].desc.flags & RF_F_BBOX) d.objs_i = d.objs_o = 1;
        d.calls = 1;
        d.prims_i = s.draws[i].desc.n_prims; d.verts_i = s.draws[i].desc.n_verts;
        d.prims_o = ds[i].prims_o; d.verts_o = 3 * ds[i].prims_o;
        d.frags_i = ds[i].frags_i; d.frags_o = ds[i].frags_o;
        c->accum.calls += 1;
        c->accum.prims_i += d.prims_i; c->accum.prims_o += d.prims_o;
        c->accum.verts_i += d.verts_i; c->accum.verts_o += d.verts_o;
        c->accum.frags_i += d.frags_i; c->accum.frags_o += d.frags_o;
        if (i + 1 == s.draws.size()) { d.time_ns = ns; c->last_draw = d; }
      }
      c->accum.time_ns += ns;
    }
    reset_slot(s);
    c->pending_dl.erase(std::remove_if(c->pending_dl.begin(), c->pending_dl.end(), [&](const rf_ctx::PendingDl& d) { return d.slot == si; }), c->pending_dl.end());
    c->flight.erase(c->flight.begin());
  }
  return result;
}

rf_status flush_impl(rf_ctx* c) {
  PassSlot& s = c->slots[c->cur];
  if (s.draws.empty() && s.clears.empty()) return RF_OK;
  const int si = c->cur;
  rf_status st = launch_pass(c, si);
  if (st) { reset_slot(s); return st; }
  c->flight.push_back(si);
  // pick the next collecting slot; if none is free, validate (blocks on the oldest pass)
  int next = -1;
  for (int k = 0; k < kSlots; k++) if (!c->slots[k].in_flight && c->slots[k].draws.empty() && c->slots[k].clears.empty()) { next = k; break; }
  if (next < 0) {
    st = validate_all(c, true);
    next = 0;
    for (int k = 0; k < kSlots; k++) if (!c->slots[k].in_flight) { next = k; break; }
  }
  c->cur = next;
  return st;
}

rf_status sync_impl(rf_ctx* c) {
  rf_status st = flush_impl(c);
  rf_status sv = validate_all(c);
  if (st == RF_OK) st = sv;
  if (st == RF_OK) RF_CUDA(c, cudaStreamSynchronize(c->stream));
  if (st == RF_OK) RF_CUDA(c, cudaStreamSynchronize(c->copy));
  return st;
}

bool fs_needs_tex(uint32_t fs) { return fs == RF_FS_TEX_CLAMP_LIT || fs == RF_FS_TEX_CLAMP || fs == RF_FS_TEX_REPEAT_POT || fs == RF_FS_TEX_ONCE; }
uint32_t fs_min_lanes(uint32_t fs) {
  switch (fs) {
    case RF_FS_COLOR3F: case RF_FS_COLOR3F_SRGB: case RF_FS_NORMAL_VIS: return 3;
    case RF_FS_COLOR4F: return 4;
    case RF_FS_TEX_CLAMP_LIT: return 5;
    default: return 2;
  }
}

rf_status queue_draw(rf_ctx* c, rf_target* target, const rf_draw* d, const float* vs_uniform_override) {
  if (!c || !target || !d) return fail(c, RF_E_INVALID, "null argument");
  if (target->ctx != c) return fail(c, RF_E_INVALID, "target belongs to another ctx");
  if (d->depth_sort > RF_SORT_BACK_TO_FRONT) return fail(c, RF_E_INVALID, "bad depth_sort");
  if (d->bbox_cull > 1 || (d->bbox_cull && d->vs == RF_VS_SPRITE)) return fail(c, RF_E_INVALID, "bbox_cull needs a vertex shader whose u[0..16] is the model-to-projection matrix");
  if (d->vs > RF_VS_SPRITE || d->fs > RF_FS_TEX_ONCE) return fail(c, RF_E_UNSUPPORTED_SHADER, "shader id not in the catalogue");
  if (d->n_attr_lanes > RF_MAX_ATTR_LANES || d->n_attr_lanes < fs_min_lanes(d->fs))
    return fail(c, RF_E_UNSUPPORTED_SHADER, "fragment shader %u needs >= %u varying lanes, got %u", d->fs, fs_min_lanes(d->fs), d->n_attr_lanes);
  if ((d->vs == RF_VS_SOLIDS && d->n_attr_lanes < 3) || (d->vs == RF_VS_SPRITE && d->n_attr_lanes < 2))
    return fail(c, RF_E_UNSUPPORTED_SHADER, "vertex shader %u lanes", d->vs);
  if (d->face_cull > RF_CULL_FRONT || d->depth_test > RF_DEPTH_GREATER || d->prim_kind > RF_PRIM_EDGES) return fail(c, RF_E_INVALID, "bad Context flag or primitive kind");
  if (fs_needs_tex(d->fs)) {
    if (!d->texture) return fail(c, RF_E_INVALID, "fragment shader needs a texture");
    if (d->fs == RF_FS_TEX_REPEAT_POT) {
      const uint32_t w = d->texture->w, h = d->texture->h;
      if (!w || !h || (w & (w - 1)) || (h & (h - 1))) return fail(c, RF_E_BAD_TEXTURE, "SamplerRepeatPot needs power-of-two dims, got %ux%u (render/tex.rs:230-231)", w, h);
    }
  }
  QueuedDraw q{};
  DrawDesc& D = q.desc;
  PassSlot& s = c->slots[c->cur];
  if (d->mesh) {
    if (d->verts || d->indices) return fail(c, RF_E_INVALID, "give either mesh or host pointers");
    if (d->mesh->stride < 3 + d->n_attr_lanes) return fail(c, RF_E_INVALID, "mesh stride too small");
    if (d->mesh->prim_kind != d->prim_kind) return fail(c, RF_E_INVALID, "rf_draw.prim_kind differs from the mesh's");
    D.verts = d->mesh->d_verts; D.indices = d->mesh->d_idx;
    D.vstride = d->mesh->stride; D.n_verts = d->mesh->n_verts; D.n_prims = d->mesh->n_prims;
    q.verts_off = q.idx_off = SIZE_MAX; q.direct = false;
  } else {
    if ((d->n_prims && !d->indices) || (d->n_verts && !d->verts)) return fail(c, RF_E_INVALID, "null geometry");
    if (d->vert_stride_f32 < 3 + d->n_attr_lanes) return fail(c, RF_E_INVALID, "vert_stride_f32 < 3 + n_attr_lanes");
    const size_t vb = (size_t)d->n_verts * d->vert_stride_f32 * 4, ib = (size_t)d->n_prims * (d->prim_kind == RF_PRIM_EDGES ? 8 : 12);
    // Page-locked caller memory (rf_host_alloc) is DMA'd to the device right here, on the copy-in stream, and the
    // call waits for it: the borrow ends at return, as in the reference, but without a staging memcpy. Pageable
    // memory is copied into pinned staging instead and uploaded with the pass.
    bool pinned_src = vb + ib >= (64u << 10);
    if (pinned_src) {
      cudaPointerAttributes av{}, ai{};
      pinned_src = cudaPointerGetAttributes(&av, d->verts) == cudaSuccess && av.type == cudaMemoryTypeHost &&
                   cudaPointerGetAttributes(&ai, d->indices) == cudaSuccess && ai.type == cudaMemoryTypeHost;
      cudaGetLastError();
    }
    if (pinned_src) {
      const size_t off = (s.direct_len + 15) & ~size_t(15);
      const size_t ioff = (off + vb + 15) & ~size_t(15);
      if (ioff + ib + 16 > s.d_direct.cap) {  // grow, preserving what earlier draws of this pass uploaded
        DevBuf nb;
        if (!nb.reserve(std::max<size_t>((ioff + ib + 16) * 2, 8u << 20))) return fail(c, RF_E_NOMEM, "direct geometry buffer");
        if (s.direct_len) RF_CUDA(c, cudaMemcpyAsync(nb.p, s.d_direct.p, s.direct_len, cudaMemcpyDeviceToDevice, c->copy_in));
        RF_CUDA(c, cudaStreamSynchronize(c->copy_in));
        s.d_direct.release();
        s.d_direct = nb;
      }
      uint8_t* db = static_cast<uint8_t*>(s.d_direct.p);
      if (vb) RF_CUDA(c, cudaMemcpyAsync(db + off, d->verts, vb, cudaMemcpyHostToDevice, c->copy_in));
      if (ib) RF_CUDA(c, cudaMemcpyAsync(db + ioff, d->indices, ib, cudaMemcpyHostToDevice, c->copy_in));
      if (c->geometry_async) {  // the caller promised to leave the arrays alone until the next rf_flush / rf_sync returns
        if (!s.ev_direct) RF_CUDA(c, cudaEventCreateWithFlags(&s.ev_direct, cudaEventDisableTiming));
        RF_CUDA(c, cudaEventRecord(s.ev_direct, c->copy_in));
        s.direct_async = true;
      } else {
        RF_CUDA(c, cudaStreamSynchronize(c->copy_in));
      }
      s.direct_len = ioff + ib;
      q.verts_off = off; q.idx_off = ioff; q.direct = true;
    } else {
      const size_t off = (s.geom_len + 15) & ~size_t(15);
      const size_t ioff = (off + vb + 15) & ~size_t(15);
      if (!s.geom.reserve(ioff + ib + 16)) return fail(c, RF_E_NOMEM, "pinned geometry staging");
      if (vb) std::memcpy(s.geom.p + off, d->verts, vb);
      if (ib) std::memcpy(s.geom.p + ioff, d->indices, ib);
      s.geom_len = ioff + ib;
      q.verts_off = off; q.idx_off = ioff; q.direct = false;
    }
    D.vstride = d->vert_stride_f32; D.n_verts = d->n_verts; D.n_prims = d->n_prims;
  }
  D.L = d->n_attr_lanes; D.persp_mask = d->persp_mask; D.vs = d->vs; D.fs = d->fs;
  D.prim_kind = d->prim_kind;
  D.flags = (uint32_t)d->face_cull | (uint32_t)d->depth_test << RF_F_DTEST_SHIFT | (d->color_write ? RF_F_CWRITE : 0u) | (d->depth_write ? RF_F_DWRITE : 0u) |
            (uint32_t)d->depth_sort << RF_F_DSORT_SHIFT | (d->bbox_cull ? RF_F_BBOX : 0u);
  std::memcpy(D.bbox, d->bbox, sizeof D.bbox);
  // an indexed mesh uses each vertex several times: transform it to screen space once (k_vertex) instead of once per use
  if ((uint64_t)D.n_prims * (d->prim_kind == RF_PRIM_EDGES ? 2 : 3) >= 2ull * D.n_verts) D.flags |= RF_F_SV;
  D.tex = d->texture ? d->texture->d_data : nullptr;
  D.tex_w = d->texture ? d->texture->w : 0; D.tex_h = d->texture ? d->texture->h : 0;
  std::memcpy(D.vs_u, vs_uniform_override ? vs_uniform_override : d->vs_uniform, sizeof D.vs_u);
  std::memcpy(D.fs_u, d->fs_uniform, sizeof D.fs_u);
  std::memcpy(D.vp, d->viewport, sizeof D.vp);
  q.target = target;
  // a pass addresses targets with 16 bits and prims with 29 (key = prim*8 + fan index)
  // (a running total: summing the queue here made a pass of n draws cost n^2 / 2 on the host — 447 ms for the 33,024 draws of a
  // 128-frame crates batch, profiles/r02_bench_crates_1gpu.json before the fix)
  if (s.prims_queued + D.n_prims >= (1u << 29) || s.draws.size() >= 60000) {
    rf_status st = flush_impl(c);
    if (st) return st;
    return queue_draw(c, target, d, vs_uniform_override);
  }
  s.draws.push_back(q);
  s.prims_queued += D.n_prims;
  return RF_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

uint32_t rf_abi_version(void) { return RF_ABI_VERSION; }

rf_status rf_ctx_create(int device, void* stream, rf_ctx** out) {
  if (!out) return RF_E_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return RF_E_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return RF_E_CUDA;
  cudaDeviceProp prop{};
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return RF_E_CUDA;
  if (prop.major != 10) return RF_E_CUDA;  // sm_100a SASS only; no fallback path
  rf_ctx* c = new rf_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  if (stream) c->stream = static_cast<cudaStream_t>(stream);
  else {
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return RF_E_CUDA; }
    c->own_stream = true;
  }
  if (cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->geo, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&c->side2, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_set_done[0], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&c->ev_set_done[1], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_pre, cudaEventDisableTiming) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming) != cudaSuccess) { rf_ctx_destroy(c); return RF_E_CUDA; }
  bool ok = cudaMalloc(&c->d_cstatus, sizeof(CtxStatus)) == cudaSuccess && cudaMemset(c->d_cstatus, 0xFF, sizeof(CtxStatus)) == cudaSuccess;  // RF_NO_POISON
  for (int k = 0; k < kSlots && ok; k++) {
    PassSlot& s = c->slots[k];
    ok = ok && cudaEventCreate(&s.ev_start) == cudaSuccess && cudaEventCreate(&s.ev_stop) == cudaSuccess;
    for (int e = 0; e <= RF_N_KERNELS && ok; e++) ok = cudaEventCreate(&s.ev_k[e]) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&s.ev_fork, cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&s.ev_join, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&s.ev_join2, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&s.ev_geo, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaHostAlloc(reinterpret_cast<void**>(&s.h_status), sizeof(HostStatus), cudaHostAllocDefault) == cudaSuccess;
  }
  ok = ok && cudaFuncSetAttribute(k_bin_sort_big, cudaFuncAttributeMaxDynamicSharedMemorySize, RF_SORT_BIG * 8) == cudaSuccess;
  ok = ok && cudaFuncSetAttribute(k_raster<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RasterSmem<3>::BYTES) == cudaSuccess;
  ok = ok && cudaFuncSetAttribute(k_raster<5, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RasterSmem<5>::BYTES) == cudaSuccess;
  ok = ok && cudaFuncSetAttribute(k_raster<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RasterSmem<8>::BYTES) == cudaSuccess;
  ok = ok && cudaFuncSetAttribute(k_raster<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RasterSmem<3>::BYTES) == cudaSuccess;
  ok = ok && cudaFuncSetAttribute(k_raster<5, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RasterSmem<5>::BYTES) == cudaSuccess;
  ok = ok && cudaFuncSetAttribute(k_raster<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RasterSmem<8>::BYTES) == cudaSuccess;
  if (!ok) { rf_ctx_destroy(c); return RF_E_CUDA; }
  *out = c;
  return RF_OK;
}

void rf_ctx_destroy(rf_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->geo) cudaStreamSynchronize(c->geo);
  if (c->side) cudaStreamSynchronize(c->side);
  cudaStreamSynchronize(c->stream);
  if (c->side2) cudaStreamSynchronize(c->side2);
  for (int k = 0; k < kSlots; k++) {
    PassSlot& s = c->slots[k];
    if (s.ev_direct) cudaEventDestroy(s.ev_direct);
    s.geom.release(); s.table.release(); s.h_dstats.release();
    s.d_geom.release(); s.d_direct.release(); s.d_table.release(); s.d_dstats.release(); s.d_status.release();
    if (s.h_status) cudaFreeHost(s.h_status);
    if (s.ev_start) cudaEventDestroy(s.ev_start);
    if (s.ev_stop) cudaEventDestroy(s.ev_stop);
    for (int e = 0; e <= RF_N_KERNELS; e++) if (s.ev_k[e]) cudaEventDestroy(s.ev_k[e]);
    if (s.ev_fork) cudaEventDestroy(s.ev_fork);
    if (s.ev_join) cudaEventDestroy(s.ev_join);
    if (s.ev_join2) cudaEventDestroy(s.ev_join2);
    if (s.ev_geo) cudaEventDestroy(s.ev_geo);
  }
  for (auto& a : c->sets) {
    a.cv.release(); a.sv.release(); a.stris.release(); a.smalls.release(); a.spans.release(); a.tris.release(); a.entries.release(); a.bins.release(); a.bins2.release();
    a.longlist.release(); a.ckpts.release(); a.chunks.release(); a.talllist.release(); a.ecks.release(); a.tiles.release(); a.cursors.release();
  }
  c->bounce.release(); c->h_bounce.release();
  for (int k = 0; k < 2; k++) if (c->ev_set_done[k]) cudaEventDestroy(c->ev_set_done[k]);
  if (c->ev_pre) cudaEventDestroy(c->ev_pre);
  if (c->geo) cudaStreamDestroy(c->geo);
  if (c->side2) cudaStreamDestroy(c->side2);
  for (uint32_t r = 0; r < c->pb.world; r++) if (r != c->pb.self && c->pb_ipc[r]) cudaIpcCloseMemHandle(c->pb.flags[r]);
  c->peer_flags.release();
  c->sdepth.release(); c->ord_tmp.release();
  for (int k = 0; k < 2; k++) { c->ord_k32[k].release(); c->ord_v[k].release(); c->ord_k64[k].release(); }
  if (c->d_cstatus) cudaFree(c->d_cstatus);
  if (c->side) cudaStreamDestroy(c->side);
  if (c->copy) { cudaStreamSynchronize(c->copy); cudaStreamDestroy(c->copy); }
  if (c->copy_in) cudaStreamDestroy(c->copy_in);
  if (c->ev_copy) cudaEventDestroy(c->ev_copy);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

const char* rf_last_error(const rf_ctx* c) { return c ? c->err.c_str() : "no context"; }

rf_status rf_ctx_set_geometry_async(rf_ctx* c, int on) {
  if (!c) return RF_E_INVALID;
  rf_status st = flush_impl(c);
  c->geometry_async = on != 0;
  return st;
}

rf_status rf_ctx_set_row_band(rf_ctx* c, uint32_t y0, uint32_t y1) {
  if (!c || y0 > y1) return fail(c, RF_E_INVALID, "bad band");
  rf_status st = flush_impl(c);
  c->band_y0 = y0; c->band_y1 = y1;
  return st;
}

rf_status rf_target_create(rf_ctx* c, uint32_t w, uint32_t h, uint32_t fmt, int has_depth, rf_target** out) {
  if (!c || !out || !w || !h || fmt > RF_FMT_RGBA4444) return fail(c, RF_E_INVALID, "bad target arguments");
  if (w > kMaxTargetDim || h > kMaxTargetDim) return fail(c, RF_E_INVALID, "target larger than %ux%u", kMaxTargetDim, kMaxTargetDim);
  cudaSetDevice(c->device);
  rf_target* t = new rf_target{c, w, h, fmt, has_depth != 0, nullptr, nullptr, nullptr, false};
  const size_t n = (size_t)w * h;
  if (cudaMalloc(&t->d_color, n * 4) != cudaSuccess) { delete t; return fail(c, RF_E_NOMEM, "colour buffer"); }
  if (has_depth && cudaMalloc(&t->d_depth, n * 4) != cudaSuccess) { cudaFree(t->d_color); delete t; return fail(c, RF_E_NOMEM, "depth buffer"); }
  cudaMemsetAsync(t->d_color, 0, n * 4, c->stream);  // Buf2::new zero-fills (util/buf.rs:155-161)
  if (has_depth) cudaMemsetAsync(t->d_depth, 0, n * 4, c->stream);
  if (has_depth && RF_LAZY_DEPTH) {
    if (cudaMalloc(&t->d_lazy, (size_t)lazy_words(t) * 4) != cudaSuccess) { cudaFree(t->d_color); cudaFree(t->d_depth); delete t; return fail(c, RF_E_NOMEM, "lazy tile table"); }
    cudaMemsetAsync(t->d_lazy, 0, (size_t)lazy_words(t) * 4, c->stream);
  }
  *out = t;
  return RF_OK;
}

void rf_target_destroy(rf_target* t) {
  if (!t) return;
  sync_impl(t->ctx);
  if (t->dl_done) cudaEventDestroy(t->dl_done);
  for (uint32_t p = 0; p < t->n_peers; p++) if (t->peer_ipc[p]) cudaIpcCloseMemHandle(t->peer_color[p]);
  cudaFree(t->d_color);
  if (t->d_depth) cudaFree(t->d_depth);
  if (t->d_lazy) cudaFree(t->d_lazy);
  delete t;
}

rf_status rf_target_clear(rf_ctx* c, rf_target* t, const uint8_t* rgba, const float* depth_recip) {
  if (!c || !t) return fail(c, RF_E_INVALID, "null argument");
  if (t->ctx != c) return fail(c, RF_E_INVALID, "target belongs to another ctx");
  // A clear is recorded at the head of a pass so that it is ordered with, and replayed with, its draws. Only draws
  // into THIS target order against it: clears of other targets join the pass being collected (clear A, draw A,
  // clear B, draw B ... is one pass).
  {
    bool touched = false;
    for (const QueuedDraw& q : c->slots[c->cur].draws) touched = touched || q.target == t;
    if (touched) { rf_status st = flush_impl(c); if (st) return st; }
  }
  QueuedClear qc{t, rgba != nullptr, depth_recip != nullptr && t->has_depth, 0u, 0u};
  if (rgba) qc.color = pack_pixel_host(t->fmt, rgba);
  if (qc.has_depth) std::memcpy(&qc.zbits, depth_recip, 4);
  if (!qc.has_color && !qc.has_depth) return RF_OK;
  // Every clear of a pass runs in ONE k_clear_multi launch, concurrently: a second clear of the same plane of the same target
  // (no draw between them, or the pass would have been flushed above) replaces the first — Frame::clear keeps the last value.
  for (QueuedClear& e : c->slots[c->cur].clears) {
    if (e.target != t) continue;
    if (qc.has_color) { e.has_color = true; e.color = qc.color; }
    if (qc.has_depth) { e.has_depth = true; e.zbits = qc.zbits; }
    return RF_OK;
  }
  c->slots[c->cur].clears.push_back(qc);
  return RF_OK;
}

rf_status rf_target_upload_color(rf_ctx* c, rf_target* t, const void* host, size_t stride) {
  if (!c || !t || !host || stride < t->w) return fail(c, RF_E_INVALID, "bad upload arguments");
  rf_status st = sync_impl(c);
  if (st) return st;
  const int hb = host_bytes(t->fmt);
  if (hb == 4) {
    RF_CUDA(c, cudaMemcpy2DAsync(t->d_color, (size_t)t->w * 4, host, stride * 4, (size_t)t->w * 4, t->h, cudaMemcpyHostToDevice, c->stream));
  } else {
    const size_t n = (size_t)t->w * t->h;
    if (!c->bounce.reserve(n * hb)) return fail(c, RF_E_NOMEM, "bounce");
    RF_CUDA(c, cudaMemcpy2DAsync(c->bounce.p, (size_t)t->w * hb, host, stride * hb, (size_t)t->w * hb, t->h, cudaMemcpyHostToDevice, c->stream));
    k_unpack_small<<<c->sm_count * 4, 256, 0, c->stream>>>(static_cast<const uint8_t*>(c->bounce.p), t->d_color, n, hb);
  }
  RF_CUDA(c, cudaStreamSynchronize(c->stream));
  return RF_OK;
}

rf_status rf_target_download_color(rf_ctx* c, rf_target* t, void* host, size_t stride) {
  if (!c || !t || !host || stride < t->w) return fail(c, RF_E_INVALID, "bad download arguments");
  rf_status st = sync_impl(c);
  if (st) return st;
  const int hb = host_bytes(t->fmt);
  if (hb == 4) {
    cudaPointerAttributes attr{};
    const bool pinned = cudaPointerGetAttributes(&attr, host) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned) {
      RF_CUDA(c, cudaMemcpy2DAsync(host, stride * 4, t->d_color, (size_t)t->w * 4, (size_t)t->w * 4, t->h, cudaMemcpyDeviceToHost, c->stream));
    } else {  // pageable destination: DMA into the pinned bounce buffer, then one host copy
      const size_t row = (size_t)t->w * 4;
      if (!c->h_bounce.reserve(row * t->h)) return fail(c, RF_E_NOMEM, "pinned bounce");
      RF_CUDA(c, cudaMemcpyAsync(c->h_bounce.p, t->d_color, row * t->h, cudaMemcpyDeviceToHost, c->stream));
      RF_CUDA(c, cudaStreamSynchronize(c->stream));
      if (stride == t->w) std::memcpy(host, c->h_bounce.p, row * t->h);
      else for (uint32_t y = 0; y < t->h; y++) std::memcpy(static_cast<uint8_t*>(host) + (size_t)y * stride * 4, c->h_bounce.p + (size_t)y * row, row);
      return RF_OK;
    }
  } else {
    const size_t n = (size_t)t->w * t->h;
    if (!c->bounce.reserve(n * hb)) return fail(c, RF_E_NOMEM, "bounce");
    k_pack_small<<<c->sm_count * 4, 256, 0, c->stream>>>(t->d_color, static_cast<uint8_t*>(c->bounce.p), n, hb);
    RF_CUDA(c, cudaMemcpy2DAsync(host, stride * hb, c->bounce.p, (size_t)t->w * hb, (size_t)t->w * hb, t->h, cudaMemcpyDeviceToHost, c->stream));
  }
  RF_CUDA(c, cudaStreamSynchronize(c->stream));
  return RF_OK;
}

rf_status rf_target_upload_depth(rf_ctx* c, rf_target* t, const float* host, size_t stride) {
  if (!c || !t || !host || stride < t->w || !t->has_depth) return fail(c, RF_E_INVALID, "bad upload arguments");
  rf_status st = sync_impl(c);
  if (st) return st;
  lazy_materialize(c, t);  // clears the marks: the plane is memory again before it is overwritten
  RF_CUDA(c, cudaMemcpy2DAsync(t->d_depth, (size_t)t->w * 4, host, stride * 4, (size_t)t->w * 4, t->h, cudaMemcpyHostToDevice, c->stream));
  RF_CUDA(c, cudaStreamSynchronize(c->stream));
  return RF_OK;
}

rf_status rf_target_download_depth(rf_ctx* c, rf_target* t, float* host, size_t stride) {
  if (!c || !t || !host || stride < t->w || !t->has_depth) return fail(c, RF_E_INVALID, "bad download arguments");
  rf_status st = sync_impl(c);
  if (st) return st;
  lazy_materialize(c, t);
  RF_CUDA(c, cudaMemcpy2DAsync(host, stride * 4, t->d_depth, (size_t)t->w * 4, (size_t)t->w * 4, t->h, cudaMemcpyDeviceToHost, c->stream));
  RF_CUDA(c, cudaStreamSynchronize(c->stream));
  return RF_OK;
}

void* rf_target_color_devptr(rf_target* t) { return t ? t->d_color : nullptr; }
void* rf_target_depth_devptr(rf_target* t) {
  if (!t) return nullptr;
  // The caller is about to read or write the plane as plain memory, with its own ordering on the ctx stream: write the lazy
  // tiles out now (stream-ordered after the passes launched so far — queued ones are flushed first) and keep the plane
  // materialised from here on.
  if (t->d_lazy && !t->lazy_off) {
    flush_impl(t->ctx);
    lazy_materialize(t->ctx, t);
    t->lazy_off = true;
  }
  return t->d_depth;
}

rf_status rf_texture_create(rf_ctx* c, uint32_t w, uint32_t h, uint32_t fmt, const void* data, size_t stride, rf_texture** out) {
  if (!c || !out || !data || !w || !h || fmt > RF_TEXEL_RGBA8888 || stride < w) return fail(c, RF_E_INVALID, "bad texture arguments");
  cudaSetDevice(c->device);
  const size_t n = (size_t)w * h;
  const int bpp = fmt == RF_TEXEL_RGB888 ? 3 : 4;
  rf_texture* t = new rf_texture{c, w, h, nullptr};
  if (cudaMalloc(&t->d_data, n * 4) != cudaSuccess) { delete t; return fail(c, RF_E_NOMEM, "texture"); }
  rf_status st = sync_impl(c);
  if (st) { cudaFree(t->d_data); delete t; return st; }
  if (bpp == 4) {
    RF_CUDA(c, cudaMemcpy2DAsync(t->d_data, (size_t)w * 4, data, stride * 4, (size_t)w * 4, h, cudaMemcpyHostToDevice, c->stream));
  } else {
    if (!c->bounce.reserve(n * 3)) return fail(c, RF_E_NOMEM, "bounce");
    RF_CUDA(c, cudaMemcpy2DAsync(c->bounce.p, (size_t)w * 3, data, stride * 3, (size_t)w * 3, h, cudaMemcpyHostToDevice, c->stream));
    k_expand_rgb<<<c->sm_count, 256, 0, c->stream>>>(static_cast<const uint8_t*>(c->bounce.p), t->d_data, n);
  }
  RF_CUDA(c, cudaStreamSynchronize(c->stream));
  *out = t;
  return RF_OK;
}

void rf_texture_destroy(rf_texture* t) {
  if (!t) return;
  sync_impl(t->ctx);
  cudaFree(t->d_data);
  delete t;
}

rf_status rf_mesh_create(rf_ctx* c, const float* verts, uint32_t n_verts, uint32_t stride, const uint32_t* indices, uint32_t n_prims, uint32_t prim_kind, rf_mesh** out) {
  if (!c || !out || !verts || !indices || stride < 3 || prim_kind > RF_PRIM_EDGES) return fail(c, RF_E_INVALID, "bad mesh arguments");
  cudaSetDevice(c->device);
  rf_mesh* m = new rf_mesh{c, nullptr, nullptr, n_verts, stride, n_prims, prim_kind};
  const size_t arity = prim_kind == RF_PRIM_EDGES ? 2 : 3;
  const size_t vb = std::max<size_t>((size_t)n_verts * stride * 4, 16), ib = std::max<size_t>((size_t)n_prims * arity * 4, 16);
  if (cudaMalloc(&m->d_verts, vb) != cudaSuccess || cudaMalloc(&m->d_idx, ib) != cudaSuccess) {
    if (m->d_verts) cudaFree(m->d_verts);
    delete m;
    return fail(c, RF_E_NOMEM, "mesh");
  }
  RF_CUDA(c, cudaMemcpyAsync(m->d_verts, verts, (size_t)n_verts * stride * 4, cudaMemcpyHostToDevice, c->stream));
  RF_CUDA(c, cudaMemcpyAsync(m->d_idx, indices, (size_t)n_prims * arity * 4, cudaMemcpyHostToDevice, c->stream));
  RF_CUDA(c, cudaStreamSynchronize(c->stream));
  *out = m;
  return RF_OK;
}

void rf_mesh_destroy(rf_mesh* m) {
  if (!m) return;
  sync_impl(m->ctx);
  cudaFree(m->d_verts);
  cudaFree(m->d_idx);
  delete m;
}

rf_status rf_render(rf_ctx* c, rf_target* target, const rf_draw* draw, rf_stats* stats_out) {
  if (c) cudaSetDevice(c->device);
  rf_status st = queue_draw(c, target, draw, nullptr);
  if (st) return st;
  if (stats_out) {
    st = sync_impl(c);
    if (st) return st;
    *stats_out = c->last_draw;
  }
  return RF_OK;
}

rf_status rf_render_frames(rf_ctx* c, rf_target* const* targets, uint32_t n_frames, const rf_draw* draw, const float* vs_uniforms) {
  if (!c || !targets || !draw || !vs_uniforms) return fail(c, RF_E_INVALID, "null argument");
  cudaSetDevice(c->device);
  for (uint32_t i = 0; i < n_frames; i++) {
    rf_status st = queue_draw(c, targets[i], draw, vs_uniforms + (size_t)i * RF_VS_UNIFORM_F32);
    if (st) return st;
  }
  return RF_OK;
}

rf_status rf_render_many(rf_ctx* c, rf_target* target, const rf_draw* draws, uint32_t n_draws) {
  if (!c || !target || (n_draws && !draws)) return fail(c, RF_E_INVALID, "null argument");
  cudaSetDevice(c->device);
  for (uint32_t i = 0; i < n_draws; i++) {
    rf_status st = queue_draw(c, target, draws + i, nullptr);
    if (st) return st;
  }
  return RF_OK;
}

rf_status rf_flush(rf_ctx* c) {
  if (!c) return RF_E_INVALID;
  cudaSetDevice(c->device);
  return flush_impl(c);
}

rf_status rf_sync(rf_ctx* c) {
  if (!c) return RF_E_INVALID;
  cudaSetDevice(c->device);
  return sync_impl(c);
}

rf_status rf_ctx_stats(rf_ctx* c, rf_stats* out, int reset) {
  if (!c || !out) return fail(c, RF_E_INVALID, "null argument");
  rf_status st = sync_impl(c);
  *out = c->accum;
  if (reset) c->accum = rf_stats{};
  return st;
}

rf_status rf_ctx_last_pass(rf_ctx* c, uint64_t* time_ns, uint32_t* n_launches) {
  if (!c) return RF_E_INVALID;
  rf_status st = sync_impl(c);
  if (time_ns) *time_ns = c->last_pass_ns;
  if (n_launches) *n_launches = c->last_pass_launches;
  return st;
}

// ---- pinned host memory for Buf2 storage: lets uploads/downloads DMA straight into caller memory
rf_status rf_host_alloc(size_t bytes, void** out) {
  if (!out || !bytes) return RF_E_INVALID;
  return cudaHostAlloc(out, bytes, cudaHostAllocDefault) == cudaSuccess ? RF_OK : RF_E_NOMEM;
}
void rf_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

// Asynchronous variant of rf_target_download_color for 4-byte formats: enqueues the copy after the
// queued draws; the pixels are in `host` after the next rf_sync(). `host` should come from rf_host_alloc.
rf_status rf_target_download_color_async(rf_ctx* c, rf_target* t, void* host, size_t stride) {
  if (!c || !t || !host || stride < t->w) return fail(c, RF_E_INVALID, "bad download arguments");
  if (host_bytes(t->fmt) != 4) return fail(c, RF_E_UNSUPPORTED, "async download needs a 4-byte pixel format");
  rf_status st = flush_impl(c);
  if (st) return st;
  // the copy runs on its own stream, after everything queued so far, and overlaps later passes
  const int last = c->flight.empty() ? -1 : c->flight.back();
  if (last >= 0) c->pending_dl.push_back(rf_ctx::PendingDl{t, host, stride, last});  // re-issued if that pass has to be replayed
  return issue_download(c, t, host, stride);
}

rf_status rf_ctx_profile(rf_ctx* c, int enable) {
  if (!c) return RF_E_INVALID;
  rf_status st = sync_impl(c);
  c->profile = enable < 0 ? 0 : (enable > 2 ? 2 : enable);
  std::memset(c->kernel_ns, 0, sizeof c->kernel_ns);
  std::memset(c->kernel_launches, 0, sizeof c->kernel_launches);
  return st;
}

rf_status rf_ctx_kernel_times(rf_ctx* c, uint64_t* ns, uint64_t* launches) {
  if (!c || !ns || !launches) return fail(c, RF_E_INVALID, "null argument");
  rf_status st = sync_impl(c);
  std::memcpy(ns, c->kernel_ns, sizeof c->kernel_ns);
  std::memcpy(launches, c->kernel_launches, sizeof c->kernel_launches);
  std::memset(c->kernel_ns, 0, sizeof c->kernel_ns);
  std::memset(c->kernel_launches, 0, sizeof c->kernel_launches);
  return st;
}

const char* rf_kernel_name(uint32_t i) {
  static const char* names[RF_N_KERNELS] = {"k_vertex", "k_assemble", "k_setup", "k_edge_ckpt", "k_walk", "k_bin_alloc", "k_bin_scatter", "k_ckpt", "k_bin_sort_warp", "k_bin_sort_big", "k_clear_untouched", "k_raster"};
  return i < RF_N_KERNELS ? names[i] : "";
}

// ---- sort-first over NVLink peer memory (rf_peer.cuh) ---------------------------------------------------------
namespace {
rf_status ensure_peer_flags(rf_ctx* c) {
  if (c->peer_flags.p) return RF_OK;
  cudaSetDevice(c->device);
  if (!c->peer_flags.reserve((RF_MAX_PEERS + 1) * RF_PEER_FLAG_STRIDE * 4)) return fail(c, RF_E_NOMEM, "peer barrier flags");
  RF_CUDA(c, cudaMemset(c->peer_flags.p, 0, c->peer_flags.cap));
  return RF_OK;
}
// resolve slot r of a peer table: an IPC handle (other process) or a raw device pointer (same process, maybe another GPU)
rf_status open_peer(rf_ctx* c, const uint8_t* ipc_handles, void* const* devptrs, uint32_t r, void** out, bool* is_ipc) {
  *is_ipc = false;
  if (devptrs) {
    *out = devptrs[r];
    if (!*out) return fail(c, RF_E_INVALID, "null peer pointer");
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, *out) == cudaSuccess && a.type == cudaMemoryTypeDevice && a.device != c->device) {
      cudaError_t e = cudaDeviceEnablePeerAccess(a.device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(c, RF_E_CUDA, "no peer access from GPU %d to GPU %d", c->device, a.device);
      cudaGetLastError();
    }
    return RF_OK;
  }
  cudaIpcMemHandle_t h;
  std::memcpy(&h, ipc_handles + (size_t)r * RF_IPC_HANDLE_BYTES, sizeof h);
  if (cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return fail(c, RF_E_CUDA, "cudaIpcOpenMemHandle failed for peer %u", r); }
  *is_ipc = true;
  return RF_OK;
}
}  // namespace

rf_status rf_ctx_peer_export(rf_ctx* c, uint8_t* ipc_handle_out, void** devptr_out) {
  if (!c) return RF_E_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == RF_IPC_HANDLE_BYTES, "handle size");
  { rf_status st = ensure_peer_flags(c); if (st) return st; }
  if (devptr_out) *devptr_out = c->peer_flags.p;
  if (ipc_handle_out) {
    cudaIpcMemHandle_t h;
    RF_CUDA(c, cudaIpcGetMemHandle(&h, c->peer_flags.p));
    std::memcpy(ipc_handle_out, &h, sizeof h);
  }
  return RF_OK;
}

rf_status rf_ctx_peer_attach(rf_ctx* c, uint32_t world, uint32_t rank, const uint8_t* ipc_handles, void* const* devptrs) {
  if (!c || world < 2 || world > RF_MAX_PEERS + 1 || rank >= world || (!ipc_handles == !devptrs)) return fail(c, RF_E_INVALID, "bad peer table");
  { rf_status st = sync_impl(c); if (st) return st; }
  { rf_status st = ensure_peer_flags(c); if (st) return st; }
  cudaSetDevice(c->device);
  PeerBarrier pb{};
  pb.world = world; pb.self = rank;
  for (uint32_t r = 0; r < world; r++) {
    if (r == rank) { pb.flags[r] = static_cast<uint32_t*>(c->peer_flags.p); continue; }
    void* p = nullptr;
    rf_status st = open_peer(c, ipc_handles, devptrs, r, &p, &c->pb_ipc[r]);
    if (st) return st;
    pb.flags[r] = static_cast<uint32_t*>(p);
  }
  c->pb = pb;
  return RF_OK;
}

rf_status rf_target_peer_export(rf_ctx* c, rf_target* t, uint8_t* ipc_handle_out, void** devptr_out) {
  if (!c || !t || t->ctx != c) return fail(c, RF_E_INVALID, "bad target");
  if (devptr_out) *devptr_out = t->d_color;
  if (ipc_handle_out) {
    cudaIpcMemHandle_t h;
    RF_CUDA(c, cudaIpcGetMemHandle(&h, t->d_color));
    std::memcpy(ipc_handle_out, &h, sizeof h);
  }
  return RF_OK;
}

rf_status rf_target_peer_attach(rf_ctx* c, rf_target* t, uint32_t world, uint32_t rank, const uint8_t* ipc_handles, void* const* devptrs) {
  if (!c || !t || t->ctx != c) return fail(c, RF_E_INVALID, "bad target");
  if (c->pb.world != world || c->pb.self != rank) return fail(c, RF_E_INVALID, "rf_ctx_peer_attach must come first, with the same world and rank");
  if (!ipc_handles == !devptrs) return fail(c, RF_E_INVALID, "give IPC handles or device pointers");
  if (t->peer_mode) return fail(c, RF_E_INVALID, "target already has peers");
  { rf_status st = sync_impl(c); if (st) return st; }
  cudaSetDevice(c->device);
  uint32_t n = 0;
  for (uint32_t r = 0; r < world; r++) {
    if (r == rank) continue;
    // an all-zero handle / null pointer: this GPU does not push into rank r (e.g. only a root rank collects the frame)
    if (devptrs ? devptrs[r] == nullptr : std::all_of(ipc_handles + (size_t)r * RF_IPC_HANDLE_BYTES, ipc_handles + (size_t)(r + 1) * RF_IPC_HANDLE_BYTES, [](uint8_t b) { return b == 0; })) continue;
    void* p = nullptr;
    rf_status st = open_peer(c, ipc_handles, devptrs, r, &p, &t->peer_ipc[n]);
    if (st) return st;
    t->peer_color[n++] = static_cast<uint32_t*>(p);
  }
  t->n_peers = n;
  t->peer_mode = true;
  return RF_OK;
}

rf_status rf_ctx_replays(rf_ctx* c, uint64_t* out) {
  if (!c || !out) return RF_E_INVALID;
  *out = c->replays;
  return RF_OK;
}

}  // extern "C"
