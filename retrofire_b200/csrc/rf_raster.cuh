// rf_raster.cuh — tile binning of triangles, per-tile submission-order sort, span checkpoints,
// and the tile rasteriser (span fill, perspective-correct varyings, texture fetch, depth test,
// colour/depth writes). Reference path: raster.rs:60-69 (fragments), target.rs:138-198.
#pragma once
#include <type_traits>

#include "rf_device.cuh"
#include "rf_geometry.cuh"

// =============================================================================================
// K3: give every non-empty tile a contiguous bin (warp prefix sum + one atomic per warp) and
// build the work lists. tile_cnt was accumulated by k_prim.
// =============================================================================================
#ifndef RF_SORT_SMALL
#define RF_SORT_SMALL 1024u   // per-warp shared-memory sort capacity (entries)
#endif
#define RF_SORT_BIG 16384u    // per-block (large smem) sort run; deeper bins are merged from runs of this size
#ifndef RF_HEAVY_BIN
#define RF_HEAVY_BIN 64u      // tiles are rasterised longest-first in three classes: >= RF_HEAVIEST_BIN, >= RF_HEAVY_BIN, rest
#endif
#ifndef RF_HEAVIEST_BIN
#define RF_HEAVIEST_BIN 160u
#endif
#define RF_SLICES 4u          // row slices of a heaviest tile (RF_TILE / RF_SLICES rows each); task word = tile | (slice+1) << 28

__global__ void __launch_bounds__(256) k_bin_alloc(PassParams P) {
  if (rf_poisoned(P)) return;
  const uint32_t lane = lane_id(), lt = (1u << lane) - 1u;
  const uint32_t n_iter = (P.n_tiles + blockDim.x * gridDim.x - 1) / (blockDim.x * gridDim.x);
  for (uint32_t it = 0; it < n_iter; it++) {
    const uint32_t t = (it * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    const uint32_t c = t < P.n_tiles ? P.tile_cnt[t] : 0u;
    const uint32_t incl = warp_scan_incl(c, lane);
    const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    const uint32_t nz = __ballot_sync(0xFFFFFFFFu, c != 0);
    const uint32_t big = __ballot_sync(0xFFFFFFFFu, c > RF_SORT_SMALL);
    const uint32_t heavy = __ballot_sync(0xFFFFFFFFu, c >= RF_HEAVY_BIN && c < RF_HEAVIEST_BIN);
    const uint32_t heaviest = __ballot_sync(0xFFFFFFFFu, c >= RF_HEAVIEST_BIN);
    uint32_t base = 0, wbase = 0, bbase = 0, hbase = 0, hhbase = 0;
    if (lane == 0 && nz) {
      base = (uint32_t)atomicAdd(&P.status->bins_needed, (unsigned long long)total);
      wbase = atomicAdd(&P.status->n_work, (uint32_t)__popc(nz));
      if (big) bbase = atomicAdd(&P.status->n_work_big, (uint32_t)__popc(big));
      if (heavy) hbase = atomicAdd(&P.status->n_work_heavy, (uint32_t)__popc(heavy));
      if (heaviest) hhbase = atomicAdd(&P.status->n_work_heaviest, (uint32_t)__popc(heaviest) * RF_SLICES);
    }
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    wbase = __shfl_sync(0xFFFFFFFFu, wbase, 0);
    bbase = __shfl_sync(0xFFFFFFFFu, bbase, 0);
    hbase = __shfl_sync(0xFFFFFFFFu, hbase, 0);
    hhbase = __shfl_sync(0xFFFFFFFFu, hhbase, 0);
    // valid entries <= entry slots <= cap_entries, so the bins always fit
    if (c != 0) {
      P.tile_off[t] = base + (incl - c);
      P.worklist[wbase + __popc(nz & lt)] = t;
      if (c > RF_SORT_SMALL) P.worklist_big[bbase + __popc(big & lt)] = t;
      // worklist_heavy holds (RF_SLICES + 1) * n_tiles words: slice tasks of the heaviest tiles in the first RF_SLICES * n_tiles
      // (every tile may be one), the heavy tiles behind them
      if (c >= RF_HEAVIEST_BIN) {  // split into RF_SLICES row slices, each rasterised by its own warp
        const uint32_t b = hhbase + __popc(heaviest & lt) * RF_SLICES;
#pragma unroll
        for (uint32_t sl = 0; sl < RF_SLICES; sl++) P.worklist_heavy[b + sl] = t | (sl + 1u) << 28;
      }
      else if (c >= RF_HEAVY_BIN) P.worklist_heavy[(size_t)RF_SLICES * P.n_tiles + hbase + __popc(heavy & lt)] = t;
      atomicMax(&P.status->max_bin, c);
    }
  }
}

// =============================================================================================
// K4: scatter the (triangle x tile) entries written by k_prim into the tile bins.
// =============================================================================================
__global__ void __launch_bounds__(256) k_bin_scatter(PassParams P) {
  if (rf_poisoned(P)) return;
  const uint32_t ne = (uint32_t)min(P.status->entries_needed, (unsigned long long)P.cap_entries);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < ne; i += gridDim.x * blockDim.x) {
    const uint4 e = __ldg(P.entries + i);
    if (e.x == RF_NO_TILE) continue;
    const uint32_t slot = P.tile_off[e.x] + atomicAdd(P.tile_fill + e.x, 1u);
    P.bins[slot] = (unsigned long long)e.y << 32 | e.z;
  }
}

// =============================================================================================
// K5: checkpoints. A span that crosses tile-column boundaries gets, for every later tile column
// it touches, the varyings at that column's first pixel — produced by the reference's own
// sequential adds (vary.rs:146-154), so tiles can start mid-span and still reproduce the running
// sums bit for bit. One thread per such span.
// =============================================================================================
template <int LT>
__global__ void __launch_bounds__(256) k_ckpt(PassParams P) {
  constexpr int SW = Rec<LT>::SW, TW = Rec<LT>::TW, KW = Rec<LT>::KW;
  constexpr int NV = 1 + LT;
  if (rf_poisoned(P)) return;
  const uint32_t nl = (uint32_t)min(P.status->long_needed, (unsigned long long)P.cap_long);
  const uint32_t lane = lane_id(), lt = (1u << lane) - 1u;
  const uint32_t n_iter = (nl + blockDim.x * gridDim.x - 1) / (blockDim.x * gridDim.x);
  for (uint32_t it = 0; it < n_iter; it++) {
    const uint32_t li = (it * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    const bool have = li < nl;
    uint32_t s = 0, own = 0, X0 = 0, n = 0, nck = 0;
    if (have) {
      const uint2 le = P.longlist[li];
      s = le.x; own = le.y;
      if (s != 0xFFFFFFFFu) {  // unused slot of a reserved block
        const uint32_t h = P.spans[(size_t)s * SW];
        X0 = h & 0xFFFFu; n = h >> 16;
        nck = ((X0 + n - 1) >> RF_TILE_SHIFT) - (X0 >> RF_TILE_SHIFT);
      }
    }
    const uint32_t incl = warp_scan_incl(nck, lane);
    const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    unsigned long long base = 0;
    if (lane == 0 && total) base = atomicAdd(&P.status->ckpts_needed, (unsigned long long)total);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (base + total > P.cap_ckpts) {
      if (lane == 0 && total) rf_overflow(P);
      continue;
    }
    if (!have || nck == 0) continue;
    const uint32_t cbase = (uint32_t)base + (incl - nck);
    uint32_t* sp = P.spans + (size_t)s * SW;
    sp[1] = cbase;
    float v[NV], dv[NV];
#pragma unroll
    for (int i = 0; i < NV; i++) v[i] = __uint_as_float(sp[2 + i]);
    const uint32_t* tr = P.tris + (size_t)(own >> 1) * TW + 8 + (own & 1u) * Rec<LT>::HS;
#pragma unroll
    for (int i = 0; i < NV; i++) dv[i] = __uint_as_float(__ldg(tr + i));
    uint32_t x = X0;
    const uint32_t xe = X0 + n;
    uint32_t k = 0;
    for (;;) {
      const uint32_t xend = ((x >> RF_TILE_SHIFT) + 1) << RF_TILE_SHIFT;
      if (xend >= xe) break;
      for (; x < xend; x++) {
#pragma unroll
        for (int i = 0; i < NV; i++) v[i] = v[i] + dv[i];
      }
      uint32_t* ck = P.ckpts + (size_t)(cbase + k) * KW;
      uint32_t w[KW];
#pragma unroll
      for (int i = 0; i < KW; i++) w[i] = i < NV ? __float_as_uint(v[i]) : 0u;
#pragma unroll
      for (int q = 0; q < KW / 2; q++) *reinterpret_cast<uint2*>(ck + 2 * q) = make_uint2(w[2 * q], w[2 * q + 1]);
      k++;
    }
  }
}

// =============================================================================================
// K6: per-tile sort of the bin by submission key (bitonic). Small bins: one WARP per tile, keys in
// that warp's shared-memory slice (<= 32 entries: registers + shuffles only). Large bins: one
// block per tile with a large shared-memory buffer.
// =============================================================================================
__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m) {
  return __shfl_xor_sync(0xFFFFFFFFu, v, m);
}

// Bitonic sort of R * 32 keys held R per lane (element r * 32 + lane in register r): the exchanges at distance < 32 are
// shuffles, those at distance >= 32 stay inside the lane. No shared memory, no bank conflicts, no warp barriers — the
// shared-memory version spent 34 % of its stall samples on conflicting accesses (profiles/r02_s5_sort_lines.txt).
template <int R>
__device__ __forceinline__ void warp_bitonic_regs(unsigned long long (&v)[R], uint32_t lane) {
#pragma unroll
  for (int k = 2; k <= R * 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
#pragma unroll
        for (int r = 0; r < R; r++) {
          const int q = r ^ (j >> 5);
          if (q > r) {
            const bool up = ((r << 5) & k) == 0;
            const unsigned long long a = v[r], b = v[q];
            if ((a > b) == up) { v[r] = b; v[q] = a; }
          }
        }
      } else {
#pragma unroll
        for (int r = 0; r < R; r++) {
          const unsigned long long o = shfl_xor_u64(v[r], j);
          const bool up = (((r << 5) | lane) & k) == 0, lower = (lane & j) == 0;
          v[r] = (up == lower) ? (o < v[r] ? o : v[r]) : (o > v[r] ? o : v[r]);
        }
      }
    }
  }
}
template <int R>
__device__ __forceinline__ void sort_bin_regs(unsigned long long* bin, uint32_t cnt, uint32_t lane) {
  unsigned long long v[R];
#pragma unroll
  for (int r = 0; r < R; r++) v[r] = (uint32_t)(r * 32) + lane < cnt ? bin[r * 32 + lane] : ~0ull;
  warp_bitonic_regs<R>(v, lane);
#pragma unroll
  for (int r = 0; r < R; r++) if ((uint32_t)(r * 32) + lane < cnt) bin[r * 32 + lane] = v[r];
}

#ifndef RF_SORT_REGS
#define RF_SORT_REGS 0   // 1: bins of 33..256 entries are sorted in registers — measured +3 % on the bunny step (78 registers, three
#endif                  // unrolled networks in the instruction cache), profiles/r02_ab_clear_placement.txt: off

#define RF_SORT_WARPS 4
__global__ void __launch_bounds__(RF_SORT_WARPS * 32) k_bin_sort_warp(PassParams P) {
  __shared__ unsigned long long sk_all[RF_SORT_WARPS][RF_SORT_SMALL];
  if (rf_poisoned(P)) return;
  const uint32_t lane = lane_id(), lt = (1u << lane) - 1u, warp = threadIdx.x >> 5;
  unsigned long long* sk = sk_all[warp];
  const uint32_t n_work = P.status->n_work;
  const uint32_t gw = blockIdx.x * RF_SORT_WARPS + warp, nw = gridDim.x * RF_SORT_WARPS;
  for (uint32_t wi = gw; wi < n_work; wi += nw) {
    const uint32_t tile = P.worklist[wi];
    const uint32_t cnt = P.tile_cnt[tile], off = P.tile_off[tile];
    if (cnt <= 1 || cnt > RF_SORT_SMALL) continue;
    unsigned long long* bin = P.bins + off;
    if (cnt <= 32) {
      unsigned long long v = lane < cnt ? bin[lane] : ~0ull;
#pragma unroll
      for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
          const unsigned long long o = shfl_xor_u64(v, j);
          const bool up = (lane & k) == 0, lower = (lane & j) == 0;
          const bool take_min = up == lower;
          v = take_min ? (o < v ? o : v) : (o > v ? o : v);
        }
      }
      if (lane < cnt) bin[lane] = v;
      continue;
    }
    if (RF_SORT_REGS && cnt <= 256) {
      if (cnt <= 64) sort_bin_regs<2>(bin, cnt, lane);
      else if (cnt <= 128) sort_bin_regs<4>(bin, cnt, lane);
      else sort_bin_regs<8>(bin, cnt, lane);
      continue;
    }
    uint32_t n2 = 64;
    while (n2 < cnt) n2 <<= 1;
    for (uint32_t i = lane; i < n2; i += 32) sk[i] = i < cnt ? bin[i] : ~0ull;
    __syncwarp();
    for (uint32_t k = 2; k <= n2; k <<= 1) {
      for (uint32_t j = k >> 1; j > 0; j >>= 1) {
        for (uint32_t p = lane; p < (n2 >> 1); p += 32) {
          const uint32_t i = ((p & ~(j - 1)) << 1) | (p & (j - 1));  // index with bit j clear
          const uint32_t l = i | j;
          const unsigned long long a = sk[i], b = sk[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { sk[i] = b; sk[l] = a; }
        }
        __syncwarp();
      }
    }
    for (uint32_t i = lane; i < cnt; i += 32) bin[i] = sk[i];
    __syncwarp();
  }
}

// bitonic sort of n2 (power of two) keys in shared memory by one block
__device__ __forceinline__ void block_bitonic(unsigned long long* sk, uint32_t n2) {
  for (uint32_t k = 2; k <= n2; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t p = threadIdx.x; p < (n2 >> 1); p += blockDim.x) {
        const uint32_t i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
        const uint32_t l = i | j;
        const unsigned long long a = sk[i], b = sk[l];
        const bool up = (i & k) == 0;
        if ((a > b) == up) { sk[i] = b; sk[l] = a; }
      }
      __syncthreads();
    }
  }
}

// Bins larger than the per-warp sorter: one block per tile. Up to RF_SORT_BIG entries are sorted in shared memory;
// deeper bins are sorted in runs of RF_SORT_BIG and the runs merged pairwise through a global scratch buffer
// (every element finds its output slot by binary search in the other run), so there is no depth limit.
__global__ void __launch_bounds__(256) k_bin_sort_big(PassParams P) {
  extern __shared__ unsigned long long skb[];
  // one decision per block (the span chain on the geo stream may set the poison while this kernel runs on the side stream)
  __shared__ uint32_t s_poisoned;
  if (threadIdx.x == 0) s_poisoned = rf_poisoned(P) ? 1u : 0u;
  __syncthreads();
  if (s_poisoned) return;
  const uint32_t n_work = P.status->n_work_big;
  for (uint32_t wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
    const uint32_t tile = P.worklist_big[wi];
    const uint32_t cnt = P.tile_cnt[tile], off = P.tile_off[tile];
    unsigned long long* bin = P.bins + off;
    for (uint32_t r0 = 0; r0 < cnt; r0 += RF_SORT_BIG) {  // sorted runs
      const uint32_t rn = min(RF_SORT_BIG, cnt - r0);
      uint32_t n2 = 1;
      while (n2 < rn) n2 <<= 1;
      __syncthreads();
      for (uint32_t i = threadIdx.x; i < n2; i += blockDim.x) skb[i] = i < rn ? bin[r0 + i] : ~0ull;
      __syncthreads();
      block_bitonic(skb, n2);
      for (uint32_t i = threadIdx.x; i < rn; i += blockDim.x) bin[r0 + i] = skb[i];
    }
    if (cnt <= RF_SORT_BIG) continue;
    unsigned long long* src = bin;
    unsigned long long* dst = P.bins2 + off;
    for (uint32_t width = RF_SORT_BIG; width < cnt; width <<= 1) {
      __syncthreads();
      for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) {
        const uint32_t pair0 = i / (2 * width) * (2 * width);
        const uint32_t a0 = pair0, a1 = min(cnt, pair0 + width), b1 = min(cnt, pair0 + 2 * width);
        const unsigned long long v = src[i];
        uint32_t lo, hi, pos;
        if (i < a1) {  // element of run A: rank among B = lower_bound
          lo = a1; hi = b1;
          while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (src[mid] < v) lo = mid + 1; else hi = mid; }
          pos = (i - a0) + (lo - a1);
        } else {       // element of run B: rank among A = upper_bound
          lo = a0; hi = a1;
          while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (src[mid] <= v) lo = mid + 1; else hi = mid; }
          pos = (i - a1) + (lo - a0);
        }
        dst[pair0 + pos] = v;
      }
      __threadfence_block();
      unsigned long long* t = src; src = dst; dst = t;
    }
    __syncthreads();
    if (src != bin)
      for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) bin[i] = src[i];
  }
}

// The first-touch clear of ONE tile without bin entries, by one warp (128-bit stores, 8 lanes per row). Under sort-first sharding
// only the rows of this GPU's band are cleared. Called by k_clear_untouched and, when the pass fuses the clear (PassParams::
// fused_clear), by the rasteriser's warps between their tiles.
__device__ __forceinline__ void clear_untouched_tile(const PassParams& P, uint32_t tile, uint32_t lane) {
  uint32_t ti = 0;
  if (P.tiles_per_target) ti = tile / P.tiles_per_target;
  else {
    uint32_t lo = 0, hi = P.n_targets;
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (P.targets[mid].tile_base <= tile) lo = mid; else hi = mid; }
    ti = lo;
  }
  const TargetDesc& T = P.targets[ti];
  const uint32_t cf = T.clear_flags;
  if (cf == 0u) return;
  const uint32_t t_w = T.w;
  const uint32_t tl = tile - T.tile_base;
  const uint32_t ty = tl / T.tiles_x, tx = tl - ty * T.tiles_x;
  const uint32_t px0 = tx << RF_TILE_SHIFT, py0 = ty << RF_TILE_SHIFT;
  const uint32_t tw = min((uint32_t)RF_TILE, t_w - px0);
  const uint32_t ya = max(py0, T.band_y0), yb = min(min(py0 + RF_TILE, T.h), T.band_y1);
  const bool vec = (t_w & 3u) == 0 && tw == RF_TILE;
  const uint32_t rsub = lane >> 3, c4 = (lane & 7u) << 2;
#pragma unroll
  for (int plane = 0; plane < 2; plane++) {
    if (!(cf & (plane ? RF_CLEAR_DEPTH : RF_CLEAR_COLOR))) continue;
    uint32_t* buf = plane ? reinterpret_cast<uint32_t*>(T.depth) : T.color;
    if (buf == nullptr) continue;
    const uint32_t val = plane ? T.clear_zbits : T.clear_color;
    if (plane && T.lazy != nullptr) {
      if (cf & RF_CLEAR_LAZY) {  // lazy depth clear: mark the tile, write nothing
        if (lane == 0) { T.lazy[tl * RF_LAZY_WORDS + 1u] = val; T.lazy[tl * RF_LAZY_WORDS] = 1u; }
        continue;
      }
      if (lane == 0) T.lazy[tl * RF_LAZY_WORDS] = 0u;  // filled as memory below: an older mark ends here
    }
    if (vec) {
      const uint4 v4 = make_uint4(val, val, val, val);
      for (uint32_t y = ya + rsub; y < yb; y += 4) *reinterpret_cast<uint4*>(buf + (size_t)y * t_w + px0 + c4) = v4;
    } else {
      for (uint32_t y = ya; y < yb; y++)
        if (lane < tw) buf[(size_t)y * t_w + px0 + lane] = val;
    }
  }
}

// =============================================================================================
// K7: tile rasteriser. One warp owns one 32x32 tile: colour and depth are staged in shared
// memory; the tile's triangles are taken in submission order, 32 at a time, and expanded into
// (triangle, row) items — one span piece per lane — each lane walking its piece pixel by pixel
// with the reference's sequential adds. Lanes whose pieces overlap on a row are serialised in
// submission order, which preserves the reference's depth-test and write semantics exactly
// (first-submitted wins ties, frags.o counts every write).
// =============================================================================================
#define RF_RASTER_WARPS 4

__device__ __forceinline__ float rust_clampf(float x, float lo, float hi) {
  if (x < lo) return lo;
  if (x > hi) return hi;
  return x;
}

// Catalogue fragment shaders. var[] already perspective-corrected. false = discard.
template <int LT>
__device__ __forceinline__ bool shade_fragment(const DrawDesc& D, uint32_t fs, const float* var, uint32_t& r, uint32_t& g, uint32_t& b, uint32_t& a) {
  a = 0xFFu;
  switch (fs) {
    case RF_FS_COLOR3F:  // color.rs:246-263
      if (LT >= 3) { r = sat_u8(256.0f * var[0]); g = sat_u8(256.0f * var[1]); b = sat_u8(256.0f * var[2]); }
      return true;
    case RF_FS_COLOR3F_SRGB:  // hello_tri.rs:18; color.rs:383-391
      if (LT >= 3) {
        r = sat_u8(256.0f * powf(var[0], 1.0f / 2.2f));
        g = sat_u8(256.0f * powf(var[1], 1.0f / 2.2f));
        b = sat_u8(256.0f * powf(var[2], 1.0f / 2.2f));
      }
      return true;
    case RF_FS_COLOR4F:  // color.rs:347-360
      if (LT >= 4) { r = sat_u8(256.0f * var[0]); g = sat_u8(256.0f * var[1]); b = sat_u8(256.0f * var[2]); a = sat_u8(256.0f * var[3]); }
      return true;
    case RF_FS_CHECKER: {  // crates.rs:33-36
      const bool eo = (var[0] > 0.5f) != (var[LT >= 2 ? 1 : 0] > 0.5f);
      r = g = b = sat_u8(256.0f * (eo ? 0.8f : 0.1f));
      return true;
    }
    case RF_FS_TEX_CLAMP_LIT: {  // crates.rs:42-47 ; tex.rs:272-304
      if (LT >= 5) {
        float ndl = 0.0f;
        ndl = ndl + var[0] * D.fs_u[0];
        ndl = ndl + var[1] * D.fs_u[1];
        ndl = ndl + var[2] * D.fs_u[2];
        ndl = fmaxf(ndl, 0.0f);
        const float kd = 0.4f + (1.0f - 0.4f) * ndl;
        const float w = (float)D.tex_w, h = (float)D.tex_h;
        const uint32_t u = sat_u32(floorf(rust_clampf(var[3] * w, 0.0f, w - 1.0f)));
        const uint32_t v = sat_u32(floorf(rust_clampf(var[4] * h, 0.0f, h - 1.0f)));
        const uint32_t c = __ldg(D.tex + (size_t)v * D.tex_w + u);
        r = sat_u8(256.0f * (((float)(c & 0xFFu) / 256.0f) * kd));
        g = sat_u8(256.0f * (((float)((c >> 8) & 0xFFu) / 256.0f) * kd));
        b = sat_u8(256.0f * (((float)((c >> 16) & 0xFFu) / 256.0f) * kd));
      }
      return true;
    }
    case RF_FS_TEX_ONCE:     // tex.rs:313-357 SamplerOnce: `(w * u) as u32`, no wrapping, no clamping; outside the texture the reference
    case RF_FS_TEX_CLAMP: {  // panics (slice index) -> the pass reports RF_E_BAD_TEXTURE.   SamplerClamp: tests/rendering.rs:30
      const float w = (float)D.tex_w, h = (float)D.tex_h;
      const float su = var[0] * w, sv = var[LT >= 2 ? 1 : 0] * h;  // Once multiplies as w * u: the product is the same
      uint32_t u = sat_u32(floorf(rust_clampf(su, 0.0f, w - 1.0f)));
      uint32_t v = sat_u32(floorf(rust_clampf(sv, 0.0f, h - 1.0f)));
      if (fs == RF_FS_TEX_ONCE) {
        u = sat_u32(su); v = sat_u32(sv);
        if (u >= D.tex_w || v >= D.tex_h) return false;  // the only way this shader returns false: the caller raises the error
      }
      const uint32_t c = __ldg(D.tex + (size_t)v * D.tex_w + u);
      r = c & 0xFFu; g = (c >> 8) & 0xFFu; b = (c >> 16) & 0xFFu; a = c >> 24;
      return true;
    }
    case RF_FS_TEX_REPEAT_POT: {  // tex.rs:218-267
      const float w = (float)D.tex_w, h = (float)D.tex_h;
      const uint32_t u = (uint32_t)sat_i32(floorf(w * var[0])) & (D.tex_w - 1);
      const uint32_t v = (uint32_t)sat_i32(floorf(h * var[LT >= 2 ? 1 : 0])) & (D.tex_h - 1);
      const uint32_t c = __ldg(D.tex + (size_t)v * D.tex_w + u);
      r = c & 0xFFu; g = (c >> 8) & 0xFFu; b = (c >> 16) & 0xFFu; a = c >> 24;
      return true;
    }
    case RF_FS_SPRITE_DISC: {  // sprites.rs:46-52
      float d2 = 0.0f;
      d2 = d2 + var[0] * var[0];
      d2 = d2 + var[LT >= 2 ? 1 : 0] * var[LT >= 2 ? 1 : 0];
      if (!(d2 < 1.0f)) return false;
      r = sat_u8(256.0f * (1.0f + (0.0f - 0.25f * d2)));
      g = sat_u8(256.0f * (1.0f + (0.0f - 0.5f * d2)));
      b = sat_u8(256.0f * (1.0f + (0.0f - 1.0f * d2)));
      return true;
    }
    default:  // RF_FS_NORMAL_VIS, curses.rs:53-56
      if (LT >= 3) {
        r = sat_u8(256.0f * (var[0] * 0.5f + 0.5f));
        g = sat_u8(256.0f * (var[1] * 0.5f + 0.5f));
        b = sat_u8(256.0f * (var[2] * 0.5f + 0.5f));
      }
      return true;
  }
}

// One fragment (target.rs:163-198): depth test -> fragment shader -> colour/depth write.
// v[0] = interpolated 1/w (the depth value), v[1..] = interpolated varyings; idx = tile-local depth index (row * RF_TILE_PITCH
// + col), gp = the pixel in the framebuffer. Returns 1 if colour was written.
// The per-warp shared-memory region (depth tile + fragment queue) addressed by an explicit 32-bit shared-window address.
// With a generic pointer ptxas re-derives the window base inside the per-fragment loops to save a register (two S2R —
// SR_CgaCtaId, SR_TID.X — plus address arithmetic per access); the volatile cvta below pins the base in a register.
#ifndef RF_GROUP_SYNC
#define RF_GROUP_SYNC 1
#endif
// L2 prefetch hint (no register result, no dependency): the records a tile's next steps will read — written by k_setup /
// k_walk long ago, 0.8 GB each on the bunny batch, so mostly out of L2 — are requested a chunk ahead of their use.
#ifndef RF_RASTER_PREFETCH
#define RF_RASTER_PREFETCH 1
#endif
#if RF_SMEM_ASM
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#else
__device__ __forceinline__ void prefetch_l2(const void*) {}
#endif

// ---- TMA bulk copies of one tile row (cp.async.bulk: SASS UBLKCP), issued per lane -------------------------------------------
#if RF_SMEM_ASM && RF_TMA_DEPTH
#define RF_TMA_ON 1
__device__ __forceinline__ void tma_mbar_init(uint32_t mbar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_mbar_expect(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void tma_load_row(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void tma_store_row(void* gdst, uint32_t smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_commit_wait_read() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tma_fence_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#else
#define RF_TMA_ON 0
#endif

struct WarpSmem {
#if RF_SMEM_ASM
  uint32_t a;
  __device__ __forceinline__ explicit WarpSmem(const void* p) {
    unsigned long long t;
    asm volatile("cvta.to.shared.u64 %0, %1;" : "=l"(t) : "l"(p));
    a = (uint32_t)t;
  }
  __device__ __forceinline__ float ldf(uint32_t w) const { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a + w * 4u) : "memory"); return v; }
  __device__ __forceinline__ uint32_t ldu(uint32_t w) const { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a + w * 4u) : "memory"); return v; }
  __device__ __forceinline__ void stf(uint32_t w, float v) const { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a + w * 4u), "f"(v) : "memory"); }
  __device__ __forceinline__ void stu(uint32_t w, uint32_t v) const { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a + w * 4u), "r"(v) : "memory"); }
  __device__ __forceinline__ void oru(uint32_t w, uint32_t v) const { asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a + w * 4u), "r"(v) : "memory"); }
  __device__ __forceinline__ void stb(uint32_t b, uint32_t v) const { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a + b), "r"(v) : "memory"); }  // byte offset
  __device__ __forceinline__ uint32_t ldb(uint32_t b) const { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a + b) : "memory"); return v; }
#else
  float* p;
  __device__ __forceinline__ explicit WarpSmem(const void* q) : p(reinterpret_cast<float*>(const_cast<void*>(q))) {}
  __device__ __forceinline__ float ldf(uint32_t w) const { return p[w]; }
  __device__ __forceinline__ uint32_t ldu(uint32_t w) const { return reinterpret_cast<const uint32_t*>(p)[w]; }
  __device__ __forceinline__ void stf(uint32_t w, float v) const { p[w] = v; }
  __device__ __forceinline__ void stu(uint32_t w, uint32_t v) const { reinterpret_cast<uint32_t*>(p)[w] = v; }
  __device__ __forceinline__ void oru(uint32_t w, uint32_t v) const { atomicOr(reinterpret_cast<uint32_t*>(p) + w, v); }
  __device__ __forceinline__ void stb(uint32_t b, uint32_t v) const { reinterpret_cast<uint8_t*>(p)[b] = (uint8_t)v; }
  __device__ __forceinline__ uint32_t ldb(uint32_t b) const { return reinterpret_cast<const uint8_t*>(p)[b]; }
#endif
};

// word (relative to the warp's shared-memory region) where a warp notes that SamplerOnce left its texture: the spare half of the
// mbarrier slot, the same offset for every lane count
#define RF_TEXERR_WORD (RF_TILE * RF_TILE_PITCH + RF_TILE + 1024u / 4u + 2u)
template <int LT>
__device__ __forceinline__ uint32_t process_fragment(const DrawDesc& D, uint32_t fs, uint32_t fmt, uint32_t sel, uint32_t* gp, WarpSmem sz, uint32_t idx, const float* v,
                                                     uint32_t pmask, uint32_t dtest, bool cwrite, bool dwrite) {
  const float z = v[0];
  if (dtest != RF_DEPTH_NONE) {  // ctx.rs:86-89: curr.partial_cmp(&new) == Some(test)
    const float curr = sz.ldf(idx);
    const bool pass = dtest == RF_DEPTH_LESS ? (curr < z) : (dtest == RF_DEPTH_EQUAL ? (curr == z) : (curr > z));
    if (!pass) return 0u;
  }
  float var[LT];
#pragma unroll
  for (int i = 0; i < LT; i++) var[i] = v[1 + i];
  if (pmask != 0) {  // Scanline::fragments: var.z_div(pos.z) on the lanes whose type divides (raster.rs:60-69)
#pragma unroll
    for (int i = 0; i < LT; i++)
      if ((pmask >> i) & 1u) var[i] = zdiv(v[1 + i], z);
  }
  uint32_t r = 0, g = 0, bl = 0, a = 0;
  if (!shade_fragment<LT>(D, fs, var, r, g, bl, a)) {  // discard: no writes at all
    // SamplerOnce never discards: false means it indexed outside its texture, where the reference panics. The warp notes it in
    // shared memory (an atomic here cost the whole kernel 330 bytes of register spills) and reports it after the tile.
    if (fs == RF_FS_TEX_ONCE) sz.stu(RF_TEXERR_WORD, 1u);
    return 0u;
  }
  // A NaN depth (0 * inf in the setup of a zero-height trapezoid half) can only be written with depth_test = None: every
  // comparison with a NaN fails (ctx.rs:86-89). The reference's x86-64 host generates the default NaN 0xFFC00000 and
  // propagates it; CUDA arithmetic generates 0x7FFFFFFF. The bits written are the host's ("NaN contract").
  if (dwrite) sz.stf(idx, z != z ? __uint_as_float(0xFFC00000u) : z);
  if (cwrite) { *gp = pack_pixel_sel(sel, fmt, r, g, bl, a); return 1u; }
  return 0u;
}

// Specialisation for the default Context (depth test Less, colour and depth writes on, ctx.rs:104-127) and a
// compile-time fragment shader / perspective mask: straight-line code, no state decoding.
template <int LT, int FS, uint32_t PMASK>
__device__ __forceinline__ uint32_t process_fragment_fixed(const DrawDesc& D, uint32_t fmt, uint32_t sel, uint32_t* gp, WarpSmem sz, uint32_t idx, const float* v) {
  const float z = v[0];
  if (!(sz.ldf(idx) < z)) return 0u;
  float var[LT];
#pragma unroll
  for (int i = 0; i < LT; i++) var[i] = ((PMASK >> i) & 1u) ? zdiv(v[1 + i], z) : v[1 + i];
  uint32_t r = 0, g = 0, bl = 0, a = 0;
  if (!shade_fragment<LT>(D, (uint32_t)FS, var, r, g, bl, a)) return 0u;
  sz.stf(idx, z);
  *gp = pack_pixel_sel(sel, fmt, r, g, bl, a);
  return 1u;
}

// Average piece length (pixels) above which a batch of pieces is walked one piece per lane (span mode);
// below it the batch is expanded to one FRAGMENT per lane, each lane doing its k sequential adds from the piece start.
// Measured (scratch/ab.sh): 6 is best at 3 varying lanes (bunny, sprites), 4 at 5 or more (crates: +5.6 %).
#ifndef RF_SPAN_MODE_MIN_AVG_3
#define RF_SPAN_MODE_MIN_AVG_3 6u
#endif
#ifndef RF_SPAN_MODE_MIN_AVG_5
#define RF_SPAN_MODE_MIN_AVG_5 4u
#endif
// Cost model: span mode runs max(pn) iterations of one fragment per lane, fragment mode runs ceil(n_frags / 32) groups, each
// paying the piece search, the shuffles and the k adds. The constants are fitted to A/B timings on the four bench workloads
// (profiles/r02_ab_span_model.txt: span mode iff max(pn) <= 7 x groups; -5 % bunny, -16 % sprites, +1.5 % small triangles).
#ifndef RF_SPAN_ITER_COST
#define RF_SPAN_ITER_COST 20u   // one span-mode iteration (0: decide by the average piece length alone)
#endif
#ifndef RF_FRAG_GROUP_COST
#define RF_FRAG_GROUP_COST 150u  // one fragment-mode group, plus RF_FRAG_K_COST per step of its k loop
#endif
#ifndef RF_FRAG_K_COST
#define RF_FRAG_K_COST 0u
#endif
template <int LT> struct RasterTune {
  static constexpr uint32_t MIN_AVG = LT == 3 ? RF_SPAN_MODE_MIN_AVG_3 : RF_SPAN_MODE_MIN_AVG_5;
};

// SMALL triangles have no span records: a (triangle, row) item takes the trapezoid half's setup {L, dl, R, dr, dv/dx} from the
// record k_assemble wrote and performs the j running-sum steps of ScanlineIter::next that lead to its row (raster.rs:84-91) —
// one ROW per lane, every lane busy (walking the rows per triangle lane ran at 21 % lane efficiency: 20 % of this kernel's
// instructions, profiles/r02_s14_raster_regions.txt). The owner table maps an item of the chunk to its triangle lane.
#define RF_OWNER_ITEMS 1024u   // items of a chunk the owner table covers (32 triangles x 32 rows); beyond that a search
template <int LT> struct RasterSmem {
  static constexpr int TILE_WORDS = RF_TILE * RF_TILE_PITCH;
  // word offsets inside a warp's region
  static constexpr int RC0 = TILE_WORDS;           // row coverage [RF_TILE]
  static constexpr int OW0 = RC0 + RF_TILE;        // owner table: one byte per item of the chunk
  static constexpr int MB0 = OW0 + (int)RF_OWNER_ITEMS / 4;  // mbarrier of the bulk loads (8 bytes), padded to keep warp regions 16-byte aligned
  static constexpr int WARP_WORDS = MB0 + 4;
  static constexpr size_t BYTES = (size_t)RF_RASTER_WARPS * WARP_WORDS * 4;
};

#ifndef RF_RASTER_MIN_BLOCKS
#define RF_RASTER_MIN_BLOCKS 6   // resident blocks per SM at 3 varying lanes; the persistent grid is this many per SM
#endif
template <int LT> struct RasterOcc { static constexpr int BLOCKS = LT == 3 ? RF_RASTER_MIN_BLOCKS : (LT == 5 ? 4 : 3); };

template <int LT, bool PEER>
__global__ void __launch_bounds__(RF_RASTER_WARPS * 32, RasterOcc<LT>::BLOCKS) k_raster(PassParams P) {
  constexpr int SW = Rec<LT>::SW, TW = Rec<LT>::TW, KW = Rec<LT>::KW, HS = Rec<LT>::HS;
  constexpr int NV = 1 + LT, NL = 2 + LT;
  using RS = RasterSmem<LT>;
  using SR = SmallRec<LT>;
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  extern __shared__ uint32_t s_raster[];
  if (rf_poisoned(P)) return;
  // A pass with a device-detected error (every check runs before this kernel) rasterises nothing, but the clears recorded
  // at its head still happen: the tiles this kernel would have touched are initialised and written back, nothing else.
  const bool clear_only = P.status->error != 0u;
  const uint32_t lane = lane_id(), lt = (1u << lane) - 1u, warp = threadIdx.x >> 5;
  // Only DEPTH is staged in shared memory: colour is write-only on this path (no blending, target.rs:187-189),
  // so passing fragments store their pixel straight to the framebuffer. __syncwarp() between dependency
  // rounds orders two writes to one pixel; untouched pixels are never read or written.
  float* sz = reinterpret_cast<float*>(s_raster + (size_t)warp * RS::WARP_WORDS);
  const WarpSmem wsm(sz);  // the same region for the per-fragment accesses: depth at [idx], coverage and the row queue behind it
  constexpr uint32_t RC0 = RS::RC0, OW0 = RS::OW0;
  // Row coverage [RF_TILE]: bit c of word r = pixel (r, c) is covered by a piece of the current fragment-mode batch. The pieces
  // OR their pixel runs in; popcount(coverage) == number of fragments <=> no two pieces of the batch share a pixel, and the
  // batch's fragment groups need no per-group same-pixel search (MATCH.ANY cost 10 % of this kernel's stall samples).
  wsm.stu(RC0 + lane, 0u);
  static_assert(RF_TEXERR_WORD == RS::MB0 + 2, "RF_TEXERR_WORD follows the RasterSmem layout");
  if (lane == 0) wsm.stu(RF_TEXERR_WORD, 0u);
#if RF_TMA_ON
#define RF_MBAR (wsm.a + RS::MB0 * 4u)  /* this warp's mbarrier for bulk loads of the depth tile */
  uint32_t mb_parity = 0;
  if (lane == 0) tma_mbar_init(RF_MBAR);
#endif
  __syncwarp();
  // fast-path selectors of the last warp-uniform draw seen (span mode | fragment mode << 4): looked up once per draw, not per batch
  uint32_t mode_draw = 0xFFFFFFFFu, mode_bits = 0;
  const uint32_t n_work = P.status->n_work, n_heaviest = P.status->n_work_heaviest, n_heavy = n_heaviest + P.status->n_work_heavy;

  // Fused first-touch clear (PassParams::fused_clear, build knob RF_FUSED_CLEAR, off): the tiles WITHOUT bin entries are filled by
  // this kernel's warps, a quota of tile ids per rasterised tile taken from a second cursor, instead of by k_clear_untouched.
  const uint32_t n_tasks = n_work + n_heavy;
  const uint32_t clear_quota = P.fused_clear ? (P.n_tiles + max(n_tasks, 1u) - 1u) / max(n_tasks, 1u) : 0u;
  auto clear_some = [&](uint32_t q) -> bool {   // false: every tile id has been handed out
    uint32_t b = 0;
    if (lane == 0) b = atomicAdd(P.cursors + 2, q);
    b = __shfl_sync(FULL, b, 0);
    if (b >= P.n_tiles) return false;
    const uint32_t e = min(b + q, P.n_tiles);
    for (uint32_t t0 = b; t0 < e; t0 += 32) {
      uint32_t m = __ballot_sync(FULL, t0 + lane < e && P.tile_cnt[t0 + lane] == 0u);
      while (m) {
        clear_untouched_tile(P, t0 + (uint32_t)__ffs(m) - 1u, lane);
        m &= m - 1u;
      }
    }
    return true;
  };

  for (;;) {
    uint32_t wi = 0;
    if (lane == 0) wi = atomicAdd(P.cursors + 1, 1u);
    wi = __shfl_sync(FULL, wi, 0);
    if (wi >= n_tasks) break;
    if (clear_quota) clear_some(clear_quota);
    const uint32_t task = wi < n_heaviest ? P.worklist_heavy[wi] : (wi < n_heavy ? P.worklist_heavy[(size_t)RF_SLICES * P.n_tiles + (wi - n_heaviest)] : P.worklist[wi - n_heavy]);
    const uint32_t tile = task & 0x0FFFFFFFu, slice = task >> 28;  // slice 0: whole tile; k+1: rows [k, k+1) * RF_TILE / RF_SLICES
    const uint32_t cnt = P.tile_cnt[tile], off = P.tile_off[tile];
    if (wi >= n_heavy && cnt >= RF_HEAVY_BIN) continue;  // already done from the heavy list
    // which target / tile coordinates
    uint32_t ti = 0;
    if (P.tiles_per_target) ti = tile / P.tiles_per_target;  // frame batch: equal targets
    else {
      uint32_t lo = 0, hi = P.n_targets;
      while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (P.targets[mid].tile_base <= tile) lo = mid; else hi = mid; }
      ti = lo;
    }
    const TargetDesc& T = P.targets[ti];
    const uint32_t tl = tile - T.tile_base;
    const uint32_t ty = tl / T.tiles_x, tx = tl - ty * T.tiles_x;
    const uint32_t px0 = tx << RF_TILE_SHIFT, py0 = ty << RF_TILE_SHIFT;
    const uint32_t tw = min((uint32_t)RF_TILE, T.w - px0), th = min((uint32_t)RF_TILE, T.h - py0);
    // read once: the shared-memory accessors are volatile asm with a memory clobber, so every later T.x would be re-loaded
    // from global memory inside the fragment loops (the format switch waited on that load: 8 % of the stall samples)
    const uint32_t t_w = T.w, t_fmt = T.fmt, t_by0 = T.band_y0, t_by1 = T.band_y1;
    const uint32_t t_sel = pack_selector(t_fmt);
    const uint32_t t_cflags = T.clear_flags;
    const bool has_depth = T.depth != nullptr;
    const bool vec = (t_w & 3u) == 0 && tw == RF_TILE;
    if (clear_only && t_cflags == 0u) continue;
    // tile rows this task owns (a heaviest tile is shared by RF_SLICES warps, each owning whole rows)
    const uint32_t r0 = slice ? min(th, (slice - 1u) * (RF_TILE / RF_SLICES)) : 0u;
    const uint32_t r1 = slice ? min(th, slice * (RF_TILE / RF_SLICES)) : th;

    uint32_t* gc = T.color + (size_t)py0 * t_w + px0;  // framebuffer address of the tile's first pixel
    // the first 32 triangles of the bin: requested before the depth tile so that the two latencies overlap
    uint32_t nb_tri = lane < cnt ? (uint32_t)P.bins[off + lane] : 0u;
    // ---- stage the depth tile. First touch after a Frame::clear of this pass: the clear value, no load at all.
    const bool depth_live = has_depth && !(clear_only && !(t_cflags & RF_CLEAR_DEPTH));
    // lazy tile (TargetDesc::lazy): an earlier first-touch clear marked it instead of filling it — start from its value, no load
    uint32_t* const lz = depth_live && T.lazy != nullptr ? T.lazy + (size_t)tl * RF_LAZY_WORDS : nullptr;
    const bool was_lazy = lz != nullptr && !(t_cflags & RF_CLEAR_DEPTH) && lz[0] != 0u;
    if (depth_live && ((t_cflags & RF_CLEAR_DEPTH) || was_lazy)) {
      const float cz = __uint_as_float(was_lazy ? lz[1] : T.clear_zbits);
      for (uint32_t r = r0; r < r1; r++) sz[r * RF_TILE_PITCH + lane] = cz;
    } else if (depth_live) {
      const float* t_depth = T.depth;
#if RF_TMA_ON
      if (vec) {  // TMA: every lane asks for one 128-byte row; the warp's mbarrier counts the bytes in
        tma_fence_proxy();  // earlier generic-proxy accesses to the tile region precede the async-proxy writes
        if (lane == 0) tma_mbar_expect(RF_MBAR, (r1 - r0) * (RF_TILE * 4u));
        __syncwarp();
        if (lane >= r0 && lane < r1) tma_load_row(wsm.a + lane * (RF_TILE_PITCH * 4u), t_depth + (size_t)(py0 + lane) * t_w + px0, RF_TILE * 4u, RF_MBAR);
        if (r1 > r0) { tma_mbar_wait(RF_MBAR, mb_parity); mb_parity ^= 1u; }
      } else
#endif
      // 128-bit coalesced loads, 8 lanes per row, 4 rows per instruction
      if (vec && r1 - r0 == RF_TILE) {  // whole tile: all eight loads in flight before the first store
        const uint32_t rsub = lane >> 3, c4 = (lane & 7u) << 2;
        float4 z[RF_TILE / 4];
#pragma unroll
        for (int i = 0; i < RF_TILE / 4; i++) z[i] = *reinterpret_cast<const float4*>(t_depth + (size_t)(py0 + rsub + 4 * i) * t_w + px0 + c4);
#pragma unroll
        for (int i = 0; i < RF_TILE / 4; i++) {
          float* e = sz + (rsub + 4 * i) * RF_TILE_PITCH + c4;
          e[0] = z[i].x; e[1] = z[i].y; e[2] = z[i].z; e[3] = z[i].w;
        }
      } else if (vec) {
        const uint32_t rsub = lane >> 3, c4 = (lane & 7u) << 2;
        for (uint32_t r = r0 + rsub; r < r1; r += 4) {
          const float4 z = *reinterpret_cast<const float4*>(t_depth + (size_t)(py0 + r) * t_w + px0 + c4);
          float* e = sz + r * RF_TILE_PITCH + c4;
          e[0] = z.x; e[1] = z.y; e[2] = z.z; e[3] = z.w;
        }
      } else {
        for (uint32_t r = r0; r < r1; r++)
          if (lane < tw) sz[r * RF_TILE_PITCH + lane] = t_depth[(size_t)(py0 + r) * t_w + px0 + lane];
      }
    }
    // ---- first touch of the colour tile: the clear colour goes to the framebuffer before the first fragment (the fragments'
    // stores to the same lines follow within microseconds, so L2 merges them: DRAM sees each line once)
    if (t_cflags & RF_CLEAR_COLOR) {
      const uint32_t cc = T.clear_color;
      if (vec) {
        const uint32_t rsub = lane >> 3, c4 = (lane & 7u) << 2;
        const uint4 cc4 = make_uint4(cc, cc, cc, cc);
        for (uint32_t r = r0 + rsub; r < r1; r += 4) *reinterpret_cast<uint4*>(gc + (size_t)r * t_w + c4) = cc4;
      } else {
        for (uint32_t r = r0; r < r1; r++)
          if (lane < tw) gc[(size_t)r * t_w + lane] = cc;
      }
    }
    __syncwarp();

    uint32_t acc_draw = 0xFFFFFFFFu;  // warp-uniform draw id of the pending frags.o / frags.i partial sums
    uint32_t acc_o = 0, acc_i = 0;    // per-lane partials
    auto flush_acc = [&]() {
      if (acc_draw != 0xFFFFFFFFu) {
        const uint32_t so = __reduce_add_sync(FULL, acc_o), si = __reduce_add_sync(FULL, acc_i);
        if (lane == 0 && so) atomicAdd(&P.dstats[acc_draw].frags_o, (unsigned long long)so);
        if (lane == 0 && si) atomicAdd(&P.dstats[acc_draw].frags_i, (unsigned long long)si);
      }
      acc_o = 0; acc_i = 0;
    };

    for (uint32_t c0 = 0; c0 < cnt && !clear_only; c0 += 32) {
      // ---- lane t: one triangle of this chunk (sorted by submission key)
      const bool t_have = c0 + lane < cnt;
      const uint32_t t_ref = nb_tri;
      nb_tri = c0 + 32 + lane < cnt ? (uint32_t)P.bins[off + c0 + 32 + lane] : 0u;  // next chunk's triangles, one chunk ahead
      const bool t_small = t_have && (t_ref & RF_BIN_SMALL) != 0u;
      const uint32_t t_tri = t_ref & ~RF_BIN_SMALL;
      // t_Y0: the first scanline's framebuffer row — signed for the other kind: scanlines above the target are drawn at row 0
      // (rows_to_scanlines); t_j0: the first scanline of this tile (or slice), t_rows: how many
      uint32_t t_sbase = 0, t_Y0 = 0, t_nU = 0, t_nrows = 0, t_j0 = 0, t_rows = 0, t_draw = 0;
      if (t_have) {
        if (t_small) {
          const uint4 h = __ldg(reinterpret_cast<const uint4*>(P.smalls + (size_t)t_tri * SR::W));  // key, draw, Y0, nU | nL << 16
          t_draw = h.y; t_Y0 = h.z; t_nU = h.w & 0xFFFFu; t_nrows = t_nU + (h.w >> 16);
        } else {
          const uint32_t* tr = P.tris + (size_t)t_tri * TW;
          const uint4 h0 = __ldg(reinterpret_cast<const uint4*>(tr));
          const uint2 h1 = __ldg(reinterpret_cast<const uint2*>(tr + 4));  // nU, nL | target << 16
          t_draw = h0.y; t_sbase = h0.z; t_Y0 = h0.w; t_nU = h1.x;
          t_nrows = h1.x + (h1.y & 0xFFFFu);
        }
        if (py0 + r1 > py0 + r0) {
          uint32_t j1;
          rows_to_scanlines((int32_t)t_Y0, t_nrows, py0 + r0, py0 + r1 - 1u, t_j0, j1);
          t_rows = j1 - t_j0;
        }
        if (RF_RASTER_PREFETCH && t_rows && !t_small) {  // this triangle's span records of the tile's rows: 24-64 bytes each
          const char* sp0 = reinterpret_cast<const char*>(P.spans + (size_t)(t_sbase + t_j0) * SW);
          prefetch_l2(sp0);
          if (t_rows * (SW * 4u) > 128u) prefetch_l2(sp0 + 128);
          if (t_rows * (SW * 4u) > 256u) prefetch_l2(sp0 + 256);
        }
      }
      if (RF_RASTER_PREFETCH && c0 + 32 + lane < cnt) {  // the next chunk's records
        const bool nsmall = (nb_tri & RF_BIN_SMALL) != 0u;
        const char* tp = nsmall ? reinterpret_cast<const char*>(P.smalls + (size_t)(nb_tri & ~RF_BIN_SMALL) * SR::W)
                                : reinterpret_cast<const char*>(P.tris + (size_t)nb_tri * TW);
        prefetch_l2(tp);
        prefetch_l2(tp + 128);
      }
      const uint32_t t_incl = warp_scan_incl(t_rows, lane);
      const uint32_t n_items = __shfl_sync(FULL, t_incl, 31);
      const uint32_t small_mask = __ballot_sync(FULL, t_small);
      // owner table: item -> triangle lane (a byte each); chunks with more items than it holds search the prefix sums instead
      const bool use_table = n_items <= RF_OWNER_ITEMS;
      if (use_table) {
        const uint32_t i0 = t_incl - t_rows;
        for (uint32_t r = 0; r < t_rows; r++) wsm.stb(OW0 * 4u + i0 + r, lane);
      }
      __syncwarp();
      {
        const uint32_t base = 0, nround = n_items;
        // (b) the round's pieces, 32 at a time, one (triangle, row) per lane. Software pipeline: the raw words of batch n+1 (a
        // queued edge state, or the span record k_setup / k_walk wrote) and its dv/dx are requested before batch n is processed.
        struct Pre {
          bool valid, small;
          uint32_t Y, draw, aux;  // aux: SMALL: unused; otherwise the span's checkpoint base
          uint32_t w[SW];
          uint32_t dvw[NV];
        };
        auto fetch = [&](uint32_t ib, Pre& p) {
          const uint32_t item = base + ib + lane;
          p.valid = ib + lane < nround;
          // owner triangle lane: from the table, or the number of lanes whose inclusive end <= item
          uint32_t ot = 0;
          if (use_table) ot = p.valid ? wsm.ldb(OW0 * 4u + item) : 0u;
          else {
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
              const uint32_t cand = ot + step;
              const uint32_t e = __shfl_sync(FULL, t_incl, (cand - 1) & 31);
              if (cand <= 32 && e <= item) ot = cand;
            }
            ot &= 31u;
          }
          const uint32_t o_incl = __shfl_sync(FULL, t_incl, ot), o_rows = __shfl_sync(FULL, t_rows, ot);
          const uint32_t o_j0 = __shfl_sync(FULL, t_j0, ot), o_Y0 = __shfl_sync(FULL, t_Y0, ot);
          const uint32_t o_sbase = __shfl_sync(FULL, t_sbase, ot), o_nU = __shfl_sync(FULL, t_nU, ot);
          const uint32_t o_tri = __shfl_sync(FULL, t_tri, ot);
          p.draw = __shfl_sync(FULL, t_draw, ot);
          p.small = ((small_mask >> ot) & 1u) != 0u;
          p.Y = 0; p.aux = 0;
          if (p.valid) {
            const uint32_t rit = item - (o_incl - o_rows);  // scanline among the triangle's scanlines in this tile
            const uint32_t j = o_j0 + rit;
            p.Y = (uint32_t)max((int32_t)o_Y0 + (int32_t)j, 0);
            const uint32_t hh = j >= o_nU ? 1u : 0u;
            if (p.small) {  // the record is read when the item is processed (L1 / L2 hits: ~7 rows share a half)
              p.w[0] = o_tri;
              p.w[1] = (hh ? j - o_nU : j) | hh << 16;
            } else {
              const uint32_t* sp = P.spans + (size_t)(o_sbase + j) * SW;
#pragma unroll
              for (int q = 0; q < SW / 2; q++) {
                const uint2 t = __ldg(reinterpret_cast<const uint2*>(sp) + q);
                p.w[2 * q] = t.x; p.w[2 * q + 1] = t.y;
              }
              const uint32_t* dp = P.tris + (size_t)o_tri * TW + 8 + hh * HS;
#pragma unroll
              for (int i = 0; i < NV; i++) p.dvw[i] = __ldg(dp + i);
            }
          }
        };
        // chunks without span-record triangles need no software pipeline (and no register copies): their items only carry indices
        const bool pipelined = __ballot_sync(FULL, t_have && !t_small && t_rows != 0u) != 0u;
        Pre cur, nxt;
        nxt.valid = false;
        if (pipelined) fetch(0, nxt);
        for (uint32_t ib = 0; ib < nround; ib += 32) {
          if (!pipelined) fetch(ib, cur);
          else {
            cur = nxt;
            nxt.valid = false;
            if (ib + 32 < nround) fetch(ib + 32, nxt);
          }
          bool valid = cur.valid;
          uint32_t py = 32 + lane, pxs = 0, pn = 0, draw = 0;
          float v[NV], dv[NV];
#pragma unroll
          for (int i = 0; i < NV; i++) { v[i] = 0.0f; dv[i] = 0.0f; }
          // ScanlineIter::next for the rows of SMALL triangles (raster.rs:84-112): the half's setup, jj steps of the running sums down
          // both edges (the same additions in the same order as the reference's iterator), then round both ends up to pixel centres,
          // align the varyings to the first centre and clip the span to this tile's columns (a SMALL triangle lies inside the
          // target: no bounds to check). The step loop runs to the largest jj of the batch.
          const bool is_small = valid && cur.small;
          float eL[NL], eR = 0.0f;
          if (__any_sync(FULL, is_small)) {
            float dl[NL], dr = 0.0f;
            uint32_t jj = 0;
#pragma unroll
            for (int i = 0; i < NL; i++) { eL[i] = 0.0f; dl[i] = 0.0f; }
            if (is_small) {
              jj = cur.w[1] & 0xFFFFu;
              const uint32_t* rec = P.smalls + (size_t)cur.w[0] * SR::W + 4 + (cur.w[1] >> 16) * SR::HW;
              constexpr int EWD = 2 * NL + 2 + NV;  // L[NL], dl[NL], R, dr, dv[NV]: contiguous in the record
              uint32_t e[(EWD + 3) & ~3];
#pragma unroll
              for (int q = 0; q < (EWD + 3) / 4; q++) {
                const uint4 t4 = __ldg(reinterpret_cast<const uint4*>(rec) + q);
                e[4 * q] = t4.x; e[4 * q + 1] = t4.y; e[4 * q + 2] = t4.z; e[4 * q + 3] = t4.w;
              }
#pragma unroll
              for (int i = 0; i < NL; i++) { eL[i] = __uint_as_float(e[SR::O_L + i]); dl[i] = __uint_as_float(e[SR::O_DL + i]); }
              eR = __uint_as_float(e[SR::O_R]); dr = __uint_as_float(e[SR::O_DR]);
#pragma unroll
              for (int i = 0; i < NV; i++) dv[i] = __uint_as_float(e[SR::O_DV + i]);
            }
            const uint32_t maxj = __reduce_max_sync(FULL, jj);
            for (uint32_t q = 0; q < maxj; q++) {
              if (q < jj) {
#pragma unroll
                for (int i = 0; i < NL; i++) eL[i] = eL[i] + dl[i];
                eR = eR + dr;
              }
            }
          }
          if (is_small) {
            const float v0x = eL[0], x1 = eR;
            const float x0r = round_up_to_half(v0x), x1r = round_up_to_half(x1);
            const uint32_t cntp = sat_u32(x1r - x0r);
            const uint32_t X0 = sat_u32(x0r), X1 = max(sat_u32(x1r), X0);
            uint32_t nn = min(cntp, X1 - X0);
            if (cur.Y < t_by0 || cur.Y >= t_by1) nn = 0;  // not this GPU's row band
            const uint32_t xs = max(X0, px0), xe = min(X0 + nn, px0 + tw);
            if (xe > xs) {
              const float tx_ = x0r - v0x;
#pragma unroll
              for (int i = 0; i < NV; i++) { const float a = eL[1 + i]; v[i] = a + ((a + dv[i]) - a) * tx_; }
              for (uint32_t k = X0; k < xs; k++) {  // the span started in an earlier tile column: the pixels before this one
#pragma unroll
                for (int i = 0; i < NV; i++) v[i] = v[i] + dv[i];
              }
              py = cur.Y - py0; pxs = xs - px0; pn = xe - xs; draw = cur.draw;
              acc_i += pn;  // frags.i (render/stats.rs): every span pixel, passed or not; the tiles' pieces add up to X1 - X0
            } else valid = false;
          } else if (valid) {
            const uint32_t X0 = cur.w[0] & 0xFFFFu, n = cur.w[0] >> 16;
            const uint32_t xs = max(X0, px0), xe = min(X0 + n, px0 + tw);
            if (n == 0 || xs >= xe) valid = false;
            else {
              py = cur.Y - py0; pxs = xs - px0; pn = xe - xs; draw = cur.draw;
              if (xs > X0) {  // the span started in an earlier tile column: take the checkpoint at this column
                const uint32_t* ck = P.ckpts + (size_t)(cur.w[1] + (tx - (X0 >> RF_TILE_SHIFT) - 1)) * KW;
#pragma unroll
                for (int q = 0; q < KW / 2; q++) {
                  const uint2 t = __ldg(reinterpret_cast<const uint2*>(ck) + q);
                  if (2 * q < NV) v[2 * q] = __uint_as_float(t.x);
                  if (2 * q + 1 < NV) v[2 * q + 1] = __uint_as_float(t.y);
                }
              } else {
#pragma unroll
                for (int i = 0; i < NV; i++) v[i] = __uint_as_float(cur.w[2 + i]);
              }
#pragma unroll
              for (int i = 0; i < NV; i++) dv[i] = __uint_as_float(cur.dvw[i]);
            }
          }
          const uint32_t vmask = __ballot_sync(FULL, valid);
          if (vmask == 0) continue;
          // ---- frags.o / frags.i bookkeeping: flush the partial sums when the draw changes
          const uint32_t d0 = __shfl_sync(FULL, draw, __ffs(vmask) - 1);
          const bool uni = __all_sync(FULL, !valid || draw == d0);
          uint32_t my_i = 0;
          if (!uni || d0 != acc_draw) {
            my_i = (valid && cur.small) ? pn : 0u;  // this batch's share, added above, belongs to the new draw(s)
            acc_i -= my_i;
            flush_acc();
            acc_draw = uni ? d0 : 0xFFFFFFFFu;
            if (uni) { acc_i = my_i; my_i = 0; }
          }
          if (!uni && my_i) atomicAdd(&P.dstats[draw].frags_i, (unsigned long long)my_i);
          uint32_t my_o = 0;

          const uint32_t f_incl = warp_scan_incl(pn, lane);
          const uint32_t n_frags = __shfl_sync(FULL, f_incl, 31);

          // warp-uniform fast-path selectors of the batch's draw (0 / 1 = generic)
          uint32_t smode = 0, fmode = uni ? 1u : 0u;
          if (uni) {
            if (d0 != mode_draw) {
              const DrawDesc& Dq = P.draws[d0];
              const uint32_t qflags = Dq.flags, qfs = Dq.fs, qpm = Dq.persp_mask, qL = Dq.L;
              const uint32_t dflt = RF_DEPTH_LESS << RF_F_DTEST_SHIFT | RF_F_CWRITE | RF_F_DWRITE;
              uint32_t sm = 0, fm = 1;
              if ((qflags & RF_F_RASTER_STATE) == dflt) {  // default Context (ctx.rs:104-127); the other flag bits concern earlier stages
                if (qfs == RF_FS_TEX_CLAMP_LIT && qpm == 0x1Fu && LT >= 5) { sm = 4; fm = 4; }
                else if (qfs == RF_FS_COLOR3F && qpm == 0u && LT >= 3) { sm = 2; fm = 2; }
                else if (qfs == RF_FS_CHECKER && qpm == 0x3u && qL == 2u && LT == 5) sm = 5;
                else if (qfs == RF_FS_SPRITE_DISC && qpm == 0x3u) fm = 3;
              }
              mode_draw = d0; mode_bits = sm | fm << 4;
            }
            if (has_depth) { smode = mode_bits & 15u; fmode = mode_bits >> 4; }
          }

          // distinct pixels covered by the batch: every piece ORs its pixel run into the coverage word of its tile row (lane r then
          // counts row r and clears the word for the next batch); equal to the fragment count <=> no two pieces share a pixel
          if (pn) wsm.oru(RC0 + py, (0xFFFFFFFFu >> (32u - pn)) << pxs);
          __syncwarp();
          const uint32_t rcov = wsm.ldu(RC0 + lane);
          wsm.stu(RC0 + lane, 0u);
          const bool no_overlap = __reduce_add_sync(FULL, (uint32_t)__popc(rcov)) == n_frags;
          bool span_mode = n_frags >= RasterTune<LT>::MIN_AVG * (uint32_t)__popc(vmask);
          if (RF_SPAN_ITER_COST != 0u && !span_mode) {
            const uint32_t maxpn = __reduce_max_sync(FULL, pn);
            span_mode = 10u + maxpn * RF_SPAN_ITER_COST <= ((n_frags + 31u) >> 5) * (RF_FRAG_GROUP_COST + RF_FRAG_K_COST * maxpn);
          }
          if (span_mode) {
            // ================= span mode: one piece per lane, walked serially =================
            // dependencies: earlier lanes on the same row whose x-range overlaps mine
            uint32_t dep = 0;
            if (!no_overlap) {
              uint32_t m = __match_any_sync(FULL, py) & lt;
              while (__any_sync(FULL, m != 0)) {
                const int jj = m ? (__ffs(m) - 1) : (int)lane;
                const uint32_t ox = __shfl_sync(FULL, pxs, jj), on = __shfl_sync(FULL, pn, jj);
                if (m) {
                  if (pxs < ox + on && ox < pxs + pn) dep |= 1u << jj;
                  m &= m - 1;
                }
              }
            }
            const DrawDesc& D = P.draws[draw];
            uint32_t done = ~vmask;
            bool pending = valid;
            while (done != FULL) {
              const bool ready = pending && (dep & ~done) == 0;
              if (ready) {
                const uint32_t pbase = py * RF_TILE_PITCH + pxs;
                uint32_t* const gp0 = gc + (size_t)py * t_w + pxs;
                if (smode == 4) {  // default Context + FS_TEX_CLAMP_LIT (crates): straight-line fragment code
                  for (uint32_t k = 0; k < pn; k++) {
                    my_o += process_fragment_fixed<LT, RF_FS_TEX_CLAMP_LIT, 0x1Fu>(D, t_fmt, t_sel, gp0 + k, wsm, pbase + k, v);
#pragma unroll
                    for (int i = 0; i < NV; i++) v[i] = v[i] + dv[i];
                  }
                } else if (smode == 2) {
                  for (uint32_t k = 0; k < pn; k++) {
                    my_o += process_fragment_fixed<LT, RF_FS_COLOR3F, 0u>(D, t_fmt, t_sel, gp0 + k, wsm, pbase + k, v);
#pragma unroll
                    for (int i = 0; i < NV; i++) v[i] = v[i] + dv[i];
                  }
                } else if (LT == 5 && smode == 5) {  // default Context + FS_CHECKER on two perspective uv lanes (the crates floor): z, u, v only
                  for (uint32_t k = 0; k < pn; k++) {
                    my_o += process_fragment_fixed<LT, RF_FS_CHECKER, 0x3u>(D, t_fmt, t_sel, gp0 + k, wsm, pbase + k, v);
#pragma unroll
                    for (int i = 0; i < 3; i++) v[i] = v[i] + dv[i];
                  }
                } else {
                  const uint32_t flags = D.flags, pmask = D.persp_mask, fs = D.fs;
                  const uint32_t dtest = has_depth ? ((flags >> RF_F_DTEST_SHIFT) & RF_F_DTEST_MASK) : (uint32_t)RF_DEPTH_NONE;
                  const bool cwrite = (flags & RF_F_CWRITE) != 0, dwrite = has_depth && (flags & RF_F_DWRITE) != 0;
                  for (uint32_t k = 0; k < pn; k++) {
                    my_o += process_fragment<LT>(D, fs, t_fmt, t_sel, gp0 + k, wsm, pbase + k, v, pmask, dtest, cwrite, dwrite);
#pragma unroll
                    for (int i = 0; i < NV; i++) v[i] = v[i] + dv[i];  // vary.rs:146-154
                  }
                }
              }
              __syncwarp();
              done |= __ballot_sync(FULL, ready);
              if (ready) pending = false;
            }
            if (uni) acc_o += my_o;
            else if (my_o) atomicAdd(&P.dstats[draw].frags_o, (unsigned long long)my_o);
          } else {
            // ================= fragment mode: one fragment per lane =================
            // Fragment f of the batch belongs to the piece whose running pixel count covers f; its lane takes the piece's
            // start values and steps over shuffles and performs the k adds of vary.rs:146-154 that bring them to pixel k.
            const uint32_t pix0 = py * RF_TILE_PITCH + pxs, gofs0 = py * t_w + pxs;
            // When the whole batch belongs to one draw (the common case) the draw state is warp-uniform and hoisted out of
            // the loop; otherwise every fragment looks its draw up.
            auto frag_loop = [&](auto mode_tag) {
              // MODE 0: per-lane draw state; 1: warp-uniform state; 2..4: warp-uniform default state with a fixed shader
              constexpr int MODE = decltype(mode_tag)::value;
              constexpr bool UNI = MODE >= 1;
              const DrawDesc& Du = P.draws[d0];
              const uint32_t u_flags = Du.flags, u_pmask = Du.persp_mask, u_fs = Du.fs;
              const uint32_t u_dtest = has_depth ? ((u_flags >> RF_F_DTEST_SHIFT) & RF_F_DTEST_MASK) : (uint32_t)RF_DEPTH_NONE;
              const bool u_cwrite = (u_flags & RF_F_CWRITE) != 0, u_dwrite = has_depth && (u_flags & RF_F_DWRITE) != 0;
              for (uint32_t fb = 0; fb < n_frags; fb += 32) {
                const uint32_t f = fb + lane;
                const bool fvalid = f < n_frags;
                // owner piece: number of lanes whose inclusive pixel count <= f
                uint32_t oi = 0;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                  const uint32_t cand = oi + step;
                  const uint32_t e = __shfl_sync(FULL, f_incl, (cand - 1) & 31);
                  if (cand <= 32 && e <= f) oi = cand;
                }
                oi &= 31u;
                const uint32_t o_end = __shfl_sync(FULL, f_incl, oi), o_pn = __shfl_sync(FULL, pn, oi), o_pix = __shfl_sync(FULL, pix0, oi), o_gofs = __shfl_sync(FULL, gofs0, oi);
                const uint32_t k = fvalid ? f - (o_end - o_pn) : 0u;
                float fv[NV], fdv[NV];
#pragma unroll
                for (int i = 0; i < NV; i++) { fv[i] = __shfl_sync(FULL, v[i], oi); fdv[i] = __shfl_sync(FULL, dv[i], oi); }
                const uint32_t maxk = __reduce_max_sync(FULL, k);
                for (uint32_t kk = 0; kk < maxk; kk++) {
                  if (kk < k) {
#pragma unroll
                    for (int i = 0; i < NV; i++) fv[i] = fv[i] + fdv[i];
                  }
                }
                const uint32_t pix = fvalid ? o_pix + k : (0x10000u + lane);
                uint32_t* const gp = gc + (o_gofs + k);
                // same pixel, submitted before me: only searched for when two pieces of the batch overlap at all
                uint32_t earlier = 0;
                bool clean = true;
                if (!no_overlap) {
                  earlier = __match_any_sync(FULL, pix) & lt;
                  clean = __all_sync(FULL, earlier == 0);
                }
                uint32_t fdraw = d0, pmask = u_pmask, fs = u_fs, dtest = u_dtest;
                bool cwrite = u_cwrite, dwrite = u_dwrite;
                if (!UNI) {
                  fdraw = __shfl_sync(FULL, draw, oi);
                  const DrawDesc& Dl = P.draws[fdraw];
                  const uint32_t flags = Dl.flags;
                  pmask = Dl.persp_mask; fs = Dl.fs;
                  dtest = has_depth ? ((flags >> RF_F_DTEST_SHIFT) & RF_F_DTEST_MASK) : (uint32_t)RF_DEPTH_NONE;
                  cwrite = (flags & RF_F_CWRITE) != 0; dwrite = has_depth && (flags & RF_F_DWRITE) != 0;
                }
                const DrawDesc& D = UNI ? Du : P.draws[fdraw];
                uint32_t wrote = 0;
                auto one = [&]() -> uint32_t {
                  if (MODE == 2) return process_fragment_fixed<LT, RF_FS_COLOR3F, 0u>(D, t_fmt, t_sel, gp, wsm, pix, fv);
                  if (MODE == 3) return process_fragment_fixed<LT, RF_FS_SPRITE_DISC, 0x3u>(D, t_fmt, t_sel, gp, wsm, pix, fv);
                  if (MODE == 4) return process_fragment_fixed<LT, RF_FS_TEX_CLAMP_LIT, 0x1Fu>(D, t_fmt, t_sel, gp, wsm, pix, fv);
                  return process_fragment<LT>(D, fs, t_fmt, t_sel, gp, wsm, pix, fv, pmask, dtest, cwrite, dwrite);
                };
                if (clean) {
                  if (fvalid) wrote = one();
#if RF_GROUP_SYNC
                  __syncwarp();  // orders this group's depth/colour writes before the next group's accesses to the same pixels
#endif
                } else {
                  uint32_t done = ~__ballot_sync(FULL, fvalid);
                  bool pending = fvalid;
                  while (done != FULL) {
                    const bool ready = pending && (earlier & ~done) == 0;
                    if (ready) wrote = one();
                    __syncwarp();
                    done |= __ballot_sync(FULL, ready);
                    if (ready) pending = false;
                  }
                }
                if (UNI) acc_o += wrote;
                else if (wrote) atomicAdd(&P.dstats[fdraw].frags_o, 1ull);
              }
            };
            if (fmode == 0) frag_loop(std::integral_constant<int, 0>{});
            else if (fmode == 2) frag_loop(std::integral_constant<int, 2>{});
            else if (fmode == 3) frag_loop(std::integral_constant<int, 3>{});
            else if (fmode == 4) frag_loop(std::integral_constant<int, 4>{});
            else frag_loop(std::integral_constant<int, 1>{});
            __syncwarp();
          }
        }
      }
      __syncwarp();  // the owner table is rewritten by the next chunk
    }
    flush_acc();
    __syncwarp();
    if (lane == 0 && wsm.ldu(RF_TEXERR_WORD) != 0u) { atomicOr(&P.status->error, RF_ERRBIT_TEXEL_OOB); wsm.stu(RF_TEXERR_WORD, 0u); }

    // ---- write the depth tile back: 128-bit coalesced stores
    if (depth_live) {
      float* t_depth = T.depth;
#if RF_TMA_ON
      if (vec) {  // TMA: every lane stores one 128-byte row from shared memory; the region is reusable once the rows have been read
        tma_fence_proxy();  // the fragments' st.shared are visible to the async proxy
        __syncwarp();
        if (lane >= r0 && lane < r1) tma_store_row(t_depth + (size_t)(py0 + lane) * t_w + px0, wsm.a + lane * (RF_TILE_PITCH * 4u), RF_TILE * 4u);
        tma_store_commit_wait_read();
      } else
#endif
      if (vec) {
        const uint32_t rsub = lane >> 3, c4 = (lane & 7u) << 2;
        for (uint32_t r = r0 + rsub; r < r1; r += 4) {
          const float* e = sz + r * RF_TILE_PITCH + c4;
          *reinterpret_cast<float4*>(t_depth + (size_t)(py0 + r) * t_w + px0 + c4) = make_float4(e[0], e[1], e[2], e[3]);
        }
      } else {
        for (uint32_t r = r0; r < r1; r++)
          if (lane < tw) t_depth[(size_t)(py0 + r) * t_w + px0 + lane] = sz[r * RF_TILE_PITCH + lane];
      }
      // the rows this task owns are real depth values now: the tile stops being lazy once every task of the tile has written
      // (a heaviest tile is RF_SLICES tasks, and a slice that starts late must still see the mark)
      if (lz != nullptr && lane == 0) {
        if (slice == 0u) lz[0] = 0u;
        else {
          __threadfence();
          if (atomicAdd(lz + 2, 1u) == RF_SLICES - 1u) { lz[2] = 0u; lz[0] = 0u; }
        }
      }
    }
    __syncwarp();

    // ---- sort-first over peer memory (rf_peer.cuh): push the finished colour rows of this tile that lie in this
    // GPU's band into the same pixels of every peer's colour buffer — P2P stores over NVLink, whole 128-byte rows,
    // once per tile whatever the overdraw, while other warps keep rasterising. The rows were just written by this
    // warp (visible after the __syncwarp above) and are still in L2.
    if (PEER) {
      const uint32_t np = T.n_peers;
      const uint32_t ya = max(py0 + r0, T.band_y0), yb = min(py0 + r1, T.band_y1);
      if (vec) {
        const uint32_t rsub = lane >> 3, c4 = (lane & 7u) << 2;
        for (uint32_t y = ya + rsub; y < yb; y += 4) {
          const size_t eo = (size_t)y * T.w + px0 + c4;
          const uint4 v = *reinterpret_cast<const uint4*>(T.color + eo);
          for (uint32_t p = 0; p < np; p++) *reinterpret_cast<uint4*>(T.peer_color[p] + eo) = v;
        }
      } else {
        for (uint32_t y = ya; y < yb; y++)
          if (lane < tw) {
            const size_t eo = (size_t)y * T.w + px0 + lane;
            const uint32_t v = T.color[eo];
            for (uint32_t p = 0; p < np; p++) T.peer_color[p][eo] = v;
          }
      }
    }
  }
  if (P.fused_clear) while (clear_some(32u)) {}   // tile ids the quotas did not reach (all of them when the pass rasterises nothing)
}

// =============================================================================================
// Frame::clear (front/src/lib.rs:103-120) for every target cleared at the head of a pass, in one
// launch: blockIdx.y selects the buffer, 128-bit stores.
// =============================================================================================
__global__ void __launch_bounds__(256) k_clear_multi(const ClearDesc* __restrict__ cl, const CtxStatus* cs, uint32_t seq) {
  if (cs->poison <= seq) return;
  const ClearDesc c = cl[blockIdx.y];
  // scalar head up to 16-byte alignment (a row band may start at any row of an odd-width target), 128-bit body, scalar tail
  const size_t head = min((size_t)c.n, (size_t)((16u - (uint32_t)(reinterpret_cast<uintptr_t>(c.ptr) & 15u)) & 15u) >> 2);
  uint32_t* body = c.ptr + head;
  const size_t nb = c.n - head, n4 = nb >> 2;
  uint4* p4 = reinterpret_cast<uint4*>(body);
  const uint4 v4 = make_uint4(c.value, c.value, c.value, c.value);
  const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  if (tid < head) c.ptr[tid] = c.value;
  for (size_t i = tid; i < c.n_lazy; i += nth) c.lazy[i] = 0u;  // the whole depth plane is written: no tile stays lazy
  for (size_t i = tid; i < n4; i += nth) p4[i] = v4;
  for (size_t i = (n4 << 2) + tid; i < nb; i += nth) body[i] = c.value;
}

// =============================================================================================
// First-touch clear, the other half: the tiles of a cleared target that no triangle of the pass
// was binned into are filled here (one warp per tile, 128-bit stores, whole 128-byte rows); the
// touched ones are initialised by k_raster. Runs on the side stream next to k_raster: the two
// write disjoint tiles. Under sort-first sharding only the rows of this GPU's band are cleared.
// =============================================================================================
#ifndef RF_CLEAR_RUN
#define RF_CLEAR_RUN 4   // k_clear_untouched: a warp takes this many consecutive tile ids; when all are untouched tiles of one tile row its stores
#endif                   // cover 128 * RF_CLEAR_RUN contiguous bytes of a framebuffer row instead of four 128-byte pieces of four rows (1: per tile)
__global__ void __launch_bounds__(256) k_clear_untouched(PassParams P) {
  if (rf_poisoned(P)) return;
  const uint32_t lane = lane_id();
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  constexpr uint32_t G = RF_CLEAR_RUN;
  static_assert(G == 1 || G == 2 || G == 4, "RF_CLEAR_RUN: 32 lanes x 16 bytes cover at most four tiles of a row");
  for (uint32_t t0 = gw * G; t0 < P.n_tiles; t0 += nw * G) {
    const uint32_t cnt = lane < G && t0 + lane < P.n_tiles ? P.tile_cnt[t0 + lane] : 1u;
    uint32_t um = __ballot_sync(0xFFFFFFFFu, cnt == 0u);
    if (um == 0u) continue;
    if (G > 1 && um == (1u << G) - 1u) {  // the whole group: one run if it lies in one tile row of one target, full tiles only
      uint32_t ti = 0;
      if (P.tiles_per_target) ti = t0 / P.tiles_per_target;
      else {
        uint32_t lo = 0, hi = P.n_targets;
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (P.targets[mid].tile_base <= t0) lo = mid; else hi = mid; }
        ti = lo;
      }
      const TargetDesc& T = P.targets[ti];
      const uint32_t t_w = T.w, tiles_x = T.tiles_x;
      const uint32_t tl = t0 - T.tile_base;
      const uint32_t ty = tl / tiles_x, tx = tl - ty * tiles_x;
      const uint32_t n_target_tiles = tiles_x * T.tiles_y;
      if (tl + G <= n_target_tiles && tx + G <= tiles_x && ((tx + G) << RF_TILE_SHIFT) <= t_w && (t_w & 3u) == 0u) {
        const uint32_t cf = T.clear_flags;
        if (cf == 0u) continue;
        const uint32_t px0 = tx << RF_TILE_SHIFT, py0 = ty << RF_TILE_SHIFT;
        const uint32_t ya = max(py0, T.band_y0), yb = min(min(py0 + RF_TILE, T.h), T.band_y1);
        constexpr uint32_t LPR = G * 8u;  // lanes per framebuffer row of the run
        const uint32_t rsub = lane / LPR, c4 = (lane % LPR) << 2;
#pragma unroll
        for (int plane = 0; plane < 2; plane++) {
          if (!(cf & (plane ? RF_CLEAR_DEPTH : RF_CLEAR_COLOR))) continue;
          uint32_t* buf = plane ? reinterpret_cast<uint32_t*>(T.depth) : T.color;
          if (buf == nullptr) continue;
          const uint32_t val = plane ? T.clear_zbits : T.clear_color;
          if (plane && T.lazy != nullptr) {
            if (cf & RF_CLEAR_LAZY) {  // lazy depth clear: mark the G tiles, write nothing
              if (lane < G) { T.lazy[(tl + lane) * RF_LAZY_WORDS + 1u] = val; T.lazy[(tl + lane) * RF_LAZY_WORDS] = 1u; }
              continue;
            }
            if (lane < G) T.lazy[(tl + lane) * RF_LAZY_WORDS] = 0u;
          }
          const uint4 v4 = make_uint4(val, val, val, val);
          for (uint32_t y = ya + rsub; y < yb; y += 32u / LPR) *reinterpret_cast<uint4*>(buf + (size_t)y * t_w + px0 + c4) = v4;
        }
        continue;
      }
    }
    while (um) {
      clear_untouched_tile(P, t0 + (uint32_t)__ffs(um) - 1u, lane);
      um &= um - 1u;
    }
  }
}

// =============================================================================================
// Lazy depth clear: write the marked tiles of ONE target out (before a download / upload of its depth plane or before its
// device pointer is handed out). One warp per tile.
// =============================================================================================
__global__ void __launch_bounds__(256) k_lazy_materialize(float* depth, uint32_t* lazy, uint32_t w, uint32_t h, uint32_t tiles_x, uint32_t n_tiles) {
  const uint32_t lane = lane_id();
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t tl = gw; tl < n_tiles; tl += nw) {
    if (lazy[tl * RF_LAZY_WORDS] == 0u) continue;
    const float z = __uint_as_float(lazy[tl * RF_LAZY_WORDS + 1u]);
    const uint32_t ty = tl / tiles_x, tx = tl - ty * tiles_x;
    const uint32_t px0 = tx << RF_TILE_SHIFT, py0 = ty << RF_TILE_SHIFT;
    const uint32_t tw = min((uint32_t)RF_TILE, w - px0), th = min((uint32_t)RF_TILE, h - py0);
    for (uint32_t r = 0; r < th; r++)
      if (lane < tw) depth[(size_t)(py0 + r) * w + px0 + lane] = z;
    __syncwarp();
    if (lane == 0) { lazy[tl * RF_LAZY_WORDS] = 0u; lazy[tl * RF_LAZY_WORDS + 2u] = 0u; }
  }
}
