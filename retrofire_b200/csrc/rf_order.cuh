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
// The two device-wide radix sorts are CUB's (cub::DeviceRadixSort, a library call): this step only
// runs for passes that contain a depth-sorted draw, never on the default path.
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
