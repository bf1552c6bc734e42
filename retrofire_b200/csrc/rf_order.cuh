// rf_order.cuh — Context::depth_sort (render.rs:180-182, 209-219): the clipped primitives of a draw are
// sorted by Render::depth (prim.rs:21-23: (a.z + b.z + c.z) / 3.0 in clip space; the trait default
// f32::INFINITY for edges, render.rs:68-70) with f32::total_cmp before they are rasterised.
//
// Here submission order is carried by a 32-bit key per screen triangle, so the sort only has to
// REPLACE the keys of a pass by ranks:  k_order_init -> radix sort by the original key (restores
// primitive order: k_assemble appends in arbitrary order) -> k_order_keys -> STABLE radix sort by
// (draw, depth bits) -> k_order_apply writes rank r into the record at position r of the order.
// Draws without depth_sort contribute depth bits 0 and therefore keep their primitive order.
// Ties: the reference uses sort_unstable_by (unspecified order of equal depths); this path and the
// oracle both define ties as "original primitive order" (stable).
// The two device-wide sorts are the LSD radix sort below (8 bits per pass, stable, three kernels per
// pass); this step only runs for passes that contain a depth-sorted draw, never on the default path.
#pragma once
#include "rf_device.cuh"

__global__ void __launch_bounds__(256) k_order_init(PassParams P, uint32_t QW, uint32_t upper, uint32_t* __restrict__ k32, uint32_t* __restrict__ vals) {
  if (rf_poisoned(P)) return;
  const uint32_t n = (uint32_t)min(P.status->stris_needed.v, (unsigned long long)P.cap_stris);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < upper; i += gridDim.x * blockDim.x) {
    const bool live = i < n;
    k32[i] = live ? P.stris[(size_t)i * QW] : 0xFFFFFFFFu;
    vals[i] = live ? i : 0xFFFFFFFFu;
  }
}

__global__ void __launch_bounds__(256) k_order_keys(PassParams P, uint32_t QW, uint32_t upper, const uint32_t* __restrict__ vals, unsigned long long* __restrict__ k64) {
  if (rf_poisoned(P)) return;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < upper; i += gridDim.x * blockDim.x) {
    const uint32_t v = vals[i];
    unsigned long long k = ~0ull;
    if (v != 0xFFFFFFFFu) {
      const uint32_t d = P.stris[(size_t)v * QW + 1] & ~RF_STRI_LINE;
      const uint32_t mode = (P.draws[d].flags >> RF_F_DSORT_SHIFT) & RF_F_DSORT_MASK;
      uint32_t bits = 0u;
      if (mode == RF_SORT_FRONT_TO_BACK) bits = P.sdepth[v];       // z.total_cmp(&w): ascending
      else if (mode == RF_SORT_BACK_TO_FRONT) bits = ~P.sdepth[v]; // w.total_cmp(&z): descending
      k = (unsigned long long)d << 32 | bits;
    }
    k64[i] = k;
  }
}

__global__ void __launch_bounds__(256) k_order_apply(PassParams P, uint32_t QW, uint32_t upper, const uint32_t* __restrict__ vals) {
  if (rf_poisoned(P)) return;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < upper; r += gridDim.x * blockDim.x) {
    const uint32_t v = vals[r];
    if (v != 0xFFFFFFFFu) P.stris[(size_t)v * QW] = r;
  }
}

// =============================================================================================
// Stable LSD radix sort of (key, uint32 value) pairs, 8 bits per pass. One warp owns a chunk of
// RF_RSORT_CHUNK consecutive elements: k_rsort_hist counts its digits, k_rsort_scan turns the
// (digit-major, chunk-minor) counts into output offsets, k_rsort_scatter re-reads the chunk in
// order and places every element — rank among the equal digits of its group of 32 by MATCH.ANY,
// groups in order — so equal digits keep their input order.
// =============================================================================================
#define RF_RSORT_CHUNK 2048u

template <class K>
__global__ void __launch_bounds__(32) k_rsort_hist(PassParams P, const K* __restrict__ keys, uint32_t n, uint32_t shift, uint32_t* __restrict__ hist, uint32_t nblk) {
  __shared__ uint32_t h[256];
  if (rf_poisoned(P)) return;
  const uint32_t lane = threadIdx.x;
  for (uint32_t d = lane; d < 256u; d += 32u) h[d] = 0u;
  __syncwarp();
  const uint32_t lo = blockIdx.x * RF_RSORT_CHUNK, hi = min(n, lo + RF_RSORT_CHUNK);
  for (uint32_t i = lo + lane; i < hi; i += 32u) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
  __syncwarp();
  for (uint32_t d = lane; d < 256u; d += 32u) hist[(size_t)d * nblk + blockIdx.x] = h[d];
}

__global__ void __launch_bounds__(256) k_rsort_scan(PassParams P, uint32_t* __restrict__ hist, uint32_t total) {
  __shared__ uint32_t part[256];
  if (rf_poisoned(P)) return;
  const uint32_t per = (total + 255u) / 256u, lo = min(total, threadIdx.x * per), hi = min(total, lo + per);
  uint32_t s = 0;
  for (uint32_t i = lo; i < hi; i++) s += hist[i];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (uint32_t t = 0; t < 256u; t++) { const uint32_t x = part[t]; part[t] = run; run += x; }
  }
  __syncthreads();
  uint32_t run = part[threadIdx.x];
  for (uint32_t i = lo; i < hi; i++) { const uint32_t x = hist[i]; hist[i] = run; run += x; }
}

template <class K>
__global__ void __launch_bounds__(32) k_rsort_scatter(PassParams P, const K* __restrict__ kin, K* __restrict__ kout, const uint32_t* __restrict__ vin,
                                                      uint32_t* __restrict__ vout, uint32_t n, uint32_t shift, const uint32_t* __restrict__ hist, uint32_t nblk) {
  __shared__ uint32_t base[256];
  if (rf_poisoned(P)) return;
  const uint32_t lane = threadIdx.x, lt = (1u << lane) - 1u;
  for (uint32_t d = lane; d < 256u; d += 32u) base[d] = hist[(size_t)d * nblk + blockIdx.x];
  __syncwarp();
  const uint32_t lo = blockIdx.x * RF_RSORT_CHUNK, hi = min(n, lo + RF_RSORT_CHUNK);
  for (uint32_t i0 = lo; i0 < hi; i0 += 32u) {
    const uint32_t i = i0 + lane;
    const bool valid = i < hi;
    const K k = valid ? kin[i] : (K)0;
    const uint32_t d = valid ? (uint32_t)(k >> shift) & 255u : 256u + lane;  // lanes past the end form groups of their own
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
    const int leader = __ffs(peers) - 1;
    uint32_t b = 0;
    if (valid && (int)lane == leader) { b = base[d]; base[d] = b + (uint32_t)__popc(peers); }
    b = __shfl_sync(0xFFFFFFFFu, b, leader);
    __syncwarp();
    if (valid) {
      const uint32_t o = b + (uint32_t)__popc(peers & lt);
      kout[o] = k;
      vout[o] = vin[i];
    }
  }
}
