// rf_raster.cuh — binning of span pieces into 32x32 tiles, per-tile submission-order sort, and
// the tile rasteriser (span fill, perspective-correct varyings, texture fetch, depth test,
// colour/depth writes). Reference path: raster.rs:60-69 (fragments), target.rs:138-198.
#pragma once
#include "rf_device.cuh"

// =============================================================================================
// K3a: count, per tile, the span pieces that fall into it. One thread per span.
// =============================================================================================
template <int LT>
__global__ void __launch_bounds__(256) k_span_count(PassParams P) {
  constexpr int SW = Rec<LT>::SW;
  if (P.cstatus->poison) return;
  const uint32_t ns = (uint32_t)min(P.status->spans_needed, (unsigned long long)P.cap_spans);
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < ns; s += gridDim.x * blockDim.x) {
    const uint4 h = __ldg(reinterpret_cast<const uint4*>(P.spans + (size_t)s * SW));
    const uint32_t n = h.z & 0xFFFFu;
    if (n == 0) continue;
    const TargetDesc& T = P.targets[h.z >> 16];
    const uint32_t trow = h.x >> RF_TILE_SHIFT;
    const uint32_t c0 = h.y >> RF_TILE_SHIFT, c1 = (h.y + n - 1) >> RF_TILE_SHIFT;
    uint32_t* cnt = P.tile_cnt + T.tile_base + trow * T.tiles_x;
    for (uint32_t c = c0; c <= c1; c++) atomicAdd(cnt + c, 1u);
  }
}

// =============================================================================================
// K3b: give every non-empty tile a contiguous bin in the piece buffer (warp prefix sum + one
// atomic per warp) and build the work lists.
// =============================================================================================
#define RF_SORT_SMALL 2048u
#define RF_SORT_BIG 16384u

__global__ void __launch_bounds__(256) k_bin_alloc(PassParams P) {
  if (P.cstatus->poison) return;
  const uint32_t lane = lane_id();
  const uint32_t n_iter = (P.n_tiles + blockDim.x * gridDim.x - 1) / (blockDim.x * gridDim.x);
  for (uint32_t it = 0; it < n_iter; it++) {
    const uint32_t t = (it * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    const uint32_t c = t < P.n_tiles ? P.tile_cnt[t] : 0u;
    const uint32_t incl = warp_scan_incl(c);
    const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    const uint32_t nz = __ballot_sync(0xFFFFFFFFu, c != 0);
    const uint32_t big = __ballot_sync(0xFFFFFFFFu, c > RF_SORT_SMALL);
    uint32_t base = 0, wbase = 0, bbase = 0;
    if (lane == 0 && nz) {
      base = (uint32_t)atomicAdd(&P.status->pieces_needed, (unsigned long long)total);
      wbase = atomicAdd(&P.status->n_work, (uint32_t)__popc(nz));
      if (big) bbase = atomicAdd(&P.status->n_work_big, (uint32_t)__popc(big));
    }
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    wbase = __shfl_sync(0xFFFFFFFFu, wbase, 0);
    bbase = __shfl_sync(0xFFFFFFFFu, bbase, 0);
    if (nz && (unsigned long long)base + total > P.cap_pieces) {
      if (lane == 0) { P.status->overflow = 1; P.cstatus->poison = 1; }
      continue;
    }
    if (c != 0) {
      P.tile_off[t] = base + (incl - c);
      P.worklist[wbase + __popc(nz & lanemask_lt())] = t;
      if (c > RF_SORT_SMALL) P.worklist_big[bbase + __popc(big & lanemask_lt())] = t;
      if (c > RF_SORT_BIG) atomicOr(&P.status->error, RF_ERRBIT_BIN_TOO_DEEP);
      atomicMax(&P.status->max_bin, c);
    }
  }
}

// =============================================================================================
// K3c: cut every span at tile-column boundaries into pieces, advancing the varyings to each
// boundary by the reference's own sequential adds (vary.rs:146-154) — this is what lets tiles
// run independently and still reproduce the running sums bit for bit. One thread per span.
// =============================================================================================
template <int LT>
__global__ void __launch_bounds__(256) k_piece_fill(PassParams P) {
  constexpr int SW = Rec<LT>::SW, HW = Rec<LT>::HW, PW = Rec<LT>::PW;
  constexpr int NV = 1 + LT;  // z + attrs
  if (P.cstatus->poison) return;
  const uint32_t ns = (uint32_t)min(P.status->spans_needed, (unsigned long long)P.cap_spans);
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < ns; s += gridDim.x * blockDim.x) {
    const uint32_t* sp = P.spans + (size_t)s * SW;
    uint32_t w[SW];
#pragma unroll
    for (int q = 0; q < SW / 4; q++) {
      const uint4 t = __ldg(reinterpret_cast<const uint4*>(sp) + q);
      w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
    }
    const uint32_t n = w[2] & 0xFFFFu;
    if (n == 0) continue;
    const uint32_t Y = w[0], X0 = w[1], half = w[3];
    const TargetDesc& T = P.targets[w[2] >> 16];
    float v[NV], dv[NV];
#pragma unroll
    for (int i = 0; i < NV; i++) v[i] = __uint_as_float(w[4 + i]);
    const uint32_t* hr = P.halves + (size_t)half * HW;
    const uint32_t key = __ldg(hr);
    const bool multi = ((X0 + n - 1) >> RF_TILE_SHIFT) != (X0 >> RF_TILE_SHIFT);
    if (multi) {
#pragma unroll
      for (int i = 0; i < NV; i++) dv[i] = __uint_as_float(__ldg(hr + 2 + i));
    }
    const uint32_t tile_row = T.tile_base + (Y >> RF_TILE_SHIFT) * T.tiles_x;
    const uint32_t xe = X0 + n;
    uint32_t x = X0;
    while (x < xe) {
      const uint32_t c = x >> RF_TILE_SHIFT;
      const uint32_t xend = min((c + 1) << RF_TILE_SHIFT, xe);
      const uint32_t m = xend - x;
      const uint32_t tile = tile_row + c;
      const uint32_t slot = P.tile_off[tile] + atomicAdd(P.tile_fill + tile, 1u);
      uint32_t o[PW];
      o[0] = key;
      o[1] = (Y & (RF_TILE - 1)) | (x & (RF_TILE - 1)) << 8 | m << 16;
      o[2] = half;
#pragma unroll
      for (int i = 0; i < NV; i++) o[3 + i] = __float_as_uint(v[i]);
#pragma unroll
      for (int i = 3 + NV; i < PW; i++) o[i] = 0u;
      uint32_t* pp = P.pieces + (size_t)slot * PW;
#pragma unroll
      for (int q = 0; q < PW / 4; q++) *reinterpret_cast<uint4*>(pp + 4 * q) = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
      if (xend < xe) {
        for (uint32_t k = 0; k < m; k++) {
#pragma unroll
          for (int i = 0; i < NV; i++) v[i] = v[i] + dv[i];
        }
      }
      x = xend;
    }
  }
}

// =============================================================================================
// K4: per-tile sort of the bin by submission key, in shared memory (bitonic). Produces, for each
// bin, the permutation `order[off + i]` = local index of the i-th piece in submission order.
// Pieces of one triangle never share a pixel, so ties between equal keys need no order.
// =============================================================================================
template <int LT, uint32_t CAP>
__global__ void __launch_bounds__(256) k_bin_sort(PassParams P, int big) {
  constexpr int PW = Rec<LT>::PW;
  extern __shared__ unsigned long long sk[];
  if (P.cstatus->poison) return;
  const uint32_t n_work = big ? P.status->n_work_big : P.status->n_work;
  const uint32_t* wl = big ? P.worklist_big : P.worklist;
  for (uint32_t wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
    const uint32_t tile = wl[wi];
    const uint32_t cnt = P.tile_cnt[tile], off = P.tile_off[tile];
    if (!big && cnt > RF_SORT_SMALL) continue;  // the big pass handles it
    if (cnt > CAP) continue;                    // flagged RF_ERRBIT_BIN_TOO_DEEP
    uint32_t n2 = 1;
    while (n2 < cnt) n2 <<= 1;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n2; i += blockDim.x)
      sk[i] = i < cnt ? ((unsigned long long)P.pieces[(size_t)(off + i) * PW] << 32 | i) : ~0ull;
    __syncthreads();
    for (uint32_t k = 2; k <= n2; k <<= 1) {
      for (uint32_t j = k >> 1; j > 0; j >>= 1) {
        for (uint32_t i = threadIdx.x; i < n2; i += blockDim.x) {
          const uint32_t l = i ^ j;
          if (l > i) {
            const unsigned long long a = sk[i], b = sk[l];
            const bool up = (i & k) == 0;
            if ((a > b) == up) { sk[i] = b; sk[l] = a; }
          }
        }
        __syncthreads();
      }
    }
    for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) P.order[off + i] = (uint32_t)sk[i];
  }
}

// =============================================================================================
// K5: tile rasteriser. One warp owns one 32x32 tile: colour and depth are staged in shared
// memory, the tile's pieces are consumed in submission order 32 at a time (one piece per lane,
// each lane walking its piece pixel by pixel with the reference's sequential adds). Lanes whose
// pieces overlap on a row are serialised in submission order, which preserves the reference's
// depth-test and write semantics exactly (first-submitted wins ties, frags.o counts every write).
// =============================================================================================
#define RF_RASTER_WARPS 4

__device__ __forceinline__ float rust_clampf(float x, float lo, float hi) {
  if (x < lo) return lo;
  if (x > hi) return hi;
  return x;
}

// Catalogue fragment shaders. var[] already perspective-corrected. false = discard.
template <int LT>
__device__ __forceinline__ bool shade_fragment(const DrawDesc& D, const float* var, uint32_t& r, uint32_t& g, uint32_t& b, uint32_t& a) {
  a = 0xFFu;
  switch (D.fs) {
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
    case RF_FS_TEX_CLAMP: {  // tests/rendering.rs:30
      const float w = (float)D.tex_w, h = (float)D.tex_h;
      const uint32_t u = sat_u32(floorf(rust_clampf(var[0] * w, 0.0f, w - 1.0f)));
      const uint32_t v = sat_u32(floorf(rust_clampf(var[LT >= 2 ? 1 : 0] * h, 0.0f, h - 1.0f)));
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

template <int LT>
__global__ void __launch_bounds__(RF_RASTER_WARPS * 32) k_raster(PassParams P) {
  constexpr int HW = Rec<LT>::HW, PW = Rec<LT>::PW;
  constexpr int NV = 1 + LT;
  __shared__ uint32_t s_color[RF_RASTER_WARPS][RF_TILE * RF_TILE_PITCH];
  __shared__ float s_depth[RF_RASTER_WARPS][RF_TILE * RF_TILE_PITCH];
  if (P.cstatus->poison || P.status->error) return;
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  uint32_t* sc = s_color[warp];
  float* sz = s_depth[warp];
  const uint32_t n_work = P.status->n_work;

  for (;;) {
    uint32_t wi = 0;
    if (lane == 0) wi = atomicAdd(P.cursors + 1, 1u);
    wi = __shfl_sync(0xFFFFFFFFu, wi, 0);
    if (wi >= n_work) break;
    const uint32_t tile = P.worklist[wi];
    const uint32_t cnt = P.tile_cnt[tile], off = P.tile_off[tile];
    // which target / tile coordinates
    uint32_t ti = 0;
    {
      uint32_t lo = 0, hi = P.n_targets;
      while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (P.targets[mid].tile_base <= tile) lo = mid; else hi = mid; }
      ti = lo;
    }
    const TargetDesc& T = P.targets[ti];
    const uint32_t tl = tile - T.tile_base;
    const uint32_t ty = tl / T.tiles_x, tx = tl - ty * T.tiles_x;
    const uint32_t px0 = tx << RF_TILE_SHIFT, py0 = ty << RF_TILE_SHIFT;
    const uint32_t tw = min((uint32_t)RF_TILE, T.w - px0), th = min((uint32_t)RF_TILE, T.h - py0);
    const bool has_depth = T.depth != nullptr;
    const bool vec = (T.w & 3u) == 0 && tw == RF_TILE;

    // ---- stage the tile: 128-bit coalesced loads, 8 lanes per row, 4 rows per instruction
    if (vec) {
      const uint32_t rsub = lane >> 3, c4 = (lane & 7u) << 2;
      for (uint32_t r = rsub; r < th; r += 4) {
        const size_t g = (size_t)(py0 + r) * T.w + px0 + c4;
        const uint4 c = *reinterpret_cast<const uint4*>(T.color + g);
        uint32_t* d = sc + r * RF_TILE_PITCH + c4;
        d[0] = c.x; d[1] = c.y; d[2] = c.z; d[3] = c.w;
        if (has_depth) {
          const float4 z = *reinterpret_cast<const float4*>(T.depth + g);
          float* e = sz + r * RF_TILE_PITCH + c4;
          e[0] = z.x; e[1] = z.y; e[2] = z.z; e[3] = z.w;
        }
      }
    } else {
      for (uint32_t r = 0; r < th; r++)
        if (lane < tw) {
          const size_t g = (size_t)(py0 + r) * T.w + px0 + lane;
          sc[r * RF_TILE_PITCH + lane] = T.color[g];
          if (has_depth) sz[r * RF_TILE_PITCH + lane] = T.depth[g];
        }
    }
    __syncwarp();

    uint32_t acc_draw = 0xFFFFFFFFu;  // warp-uniform draw id of the pending frags.o partial sums
    uint32_t acc_o = 0;               // per-lane partial

    for (uint32_t b0 = 0; b0 < cnt; b0 += 32) {
      const bool valid = b0 + lane < cnt;
      uint32_t py = 32 + lane, pxs = 0, pn = 0, draw = 0;
      float v[NV], dv[NV];
#pragma unroll
      for (int i = 0; i < NV; i++) { v[i] = 0.0f; dv[i] = 0.0f; }
      if (valid) {
        const uint32_t li = P.order[off + b0 + lane];
        const uint32_t* pp = P.pieces + (size_t)(off + li) * PW;
        uint32_t w[PW];
#pragma unroll
        for (int q = 0; q < PW / 4; q++) {
          const uint4 t = __ldg(reinterpret_cast<const uint4*>(pp) + q);
          w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
        }
        py = w[1] & 0xFFu; pxs = (w[1] >> 8) & 0xFFu; pn = w[1] >> 16;
#pragma unroll
        for (int i = 0; i < NV; i++) v[i] = __uint_as_float(w[3 + i]);
        const uint32_t* hr = P.halves + (size_t)w[2] * HW;
        uint32_t hw[HW];
#pragma unroll
        for (int q = 0; q < HW / 4; q++) {
          const uint4 t = __ldg(reinterpret_cast<const uint4*>(hr) + q);
          hw[4 * q] = t.x; hw[4 * q + 1] = t.y; hw[4 * q + 2] = t.z; hw[4 * q + 3] = t.w;
        }
        draw = hw[1];
#pragma unroll
        for (int i = 0; i < NV; i++) dv[i] = __uint_as_float(hw[2 + i]);
      }
      // ---- dependencies: earlier lanes on the same row whose x-range overlaps mine
      uint32_t dep = 0;
      {
        uint32_t m = __match_any_sync(0xFFFFFFFFu, py) & lanemask_lt();
        while (__any_sync(0xFFFFFFFFu, m != 0)) {
          const int j = m ? (__ffs(m) - 1) : (int)lane;
          const uint32_t ox = __shfl_sync(0xFFFFFFFFu, pxs, j), on = __shfl_sync(0xFFFFFFFFu, pn, j);
          if (m) {
            if (pxs < ox + on && ox < pxs + pn) dep |= 1u << j;
            m &= m - 1;
          }
        }
      }
      // ---- frags.o bookkeeping: flush partial sums when the draw changes
      const uint32_t d0 = __shfl_sync(0xFFFFFFFFu, draw, __ffs(__ballot_sync(0xFFFFFFFFu, valid)) - 1);
      const bool uni = __all_sync(0xFFFFFFFFu, !valid || draw == d0);
      if (!uni || d0 != acc_draw) {
        if (acc_draw != 0xFFFFFFFFu) {
          uint32_t s = acc_o;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
          if (lane == 0 && s) atomicAdd(&P.dstats[acc_draw].frags_o, (unsigned long long)s);
        }
        acc_o = 0;
        acc_draw = uni ? d0 : 0xFFFFFFFFu;
      }
      uint32_t my_o = 0;

      const DrawDesc& D = P.draws[draw];
      const uint32_t flags = valid ? D.flags : 0u;
      const uint32_t pmask = valid ? D.persp_mask : 0u;
      const uint32_t dtest = has_depth ? ((flags >> RF_F_DTEST_SHIFT) & RF_F_DTEST_MASK) : (uint32_t)RF_DEPTH_NONE;
      const bool cwrite = (flags & RF_F_CWRITE) != 0, dwrite = has_depth && (flags & RF_F_DWRITE) != 0;

      uint32_t done = ~__ballot_sync(0xFFFFFFFFu, valid);
      bool pending = valid;
      while (done != 0xFFFFFFFFu) {
        const bool ready = pending && (dep & ~done) == 0;
        uint32_t maxn = ready ? pn : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) maxn = max(maxn, __shfl_xor_sync(0xFFFFFFFFu, maxn, o));
        if (ready) {
          const uint32_t base = py * RF_TILE_PITCH + pxs;
          for (uint32_t k = 0; k < pn; k++) {
            const float z = v[0];
            bool pass = true;  // ctx.rs:86-89: curr.partial_cmp(&new) == Some(test)
            if (dtest != RF_DEPTH_NONE) {
              const float curr = sz[base + k];
              pass = dtest == RF_DEPTH_LESS ? (curr < z) : (dtest == RF_DEPTH_EQUAL ? (curr == z) : (curr > z));
            }
            if (pass) {
              float var[LT];
#pragma unroll
              for (int i = 0; i < LT; i++) var[i] = ((pmask >> i) & 1u) ? v[1 + i] / z : v[1 + i];  // raster.rs:60-69
              uint32_t r = 0, g = 0, bl = 0, a = 0;
              if (shade_fragment<LT>(D, var, r, g, bl, a)) {
                if (cwrite) { my_o++; sc[base + k] = pack_pixel(T.fmt, r, g, bl, a); }
                if (dwrite) sz[base + k] = z;
              }
            }
#pragma unroll
            for (int i = 0; i < NV; i++) v[i] = v[i] + dv[i];  // vary.rs:146-154
          }
        }
        (void)maxn;
        __syncwarp();
        done |= __ballot_sync(0xFFFFFFFFu, ready);
        if (ready) pending = false;
      }
      if (uni) acc_o += my_o;
      else if (my_o) atomicAdd(&P.dstats[draw].frags_o, (unsigned long long)my_o);
    }
    if (acc_draw != 0xFFFFFFFFu) {
      uint32_t s = acc_o;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
      if (lane == 0 && s) atomicAdd(&P.dstats[acc_draw].frags_o, (unsigned long long)s);
    }
    __syncwarp();

    // ---- write the tile back: 128-bit coalesced stores
    if (vec) {
      const uint32_t rsub = lane >> 3, c4 = (lane & 7u) << 2;
      for (uint32_t r = rsub; r < th; r += 4) {
        const size_t g = (size_t)(py0 + r) * T.w + px0 + c4;
        const uint32_t* d = sc + r * RF_TILE_PITCH + c4;
        *reinterpret_cast<uint4*>(T.color + g) = make_uint4(d[0], d[1], d[2], d[3]);
        if (has_depth) {
          const float* e = sz + r * RF_TILE_PITCH + c4;
          *reinterpret_cast<float4*>(T.depth + g) = make_float4(e[0], e[1], e[2], e[3]);
        }
      }
    } else {
      for (uint32_t r = 0; r < th; r++)
        if (lane < tw) {
          const size_t g = (size_t)(py0 + r) * T.w + px0 + lane;
          T.color[g] = sc[r * RF_TILE_PITCH + lane];
          if (has_depth) T.depth[g] = sz[r * RF_TILE_PITCH + lane];
        }
    }
    __syncwarp();
  }
}
