// rf_peer.cuh — sort-first sharding over NVLink peer memory: the exchange step fused into the rasteriser.
//
// Every GPU rasterises its row band and k_raster<LT, true>, as soon as it has finished a tile, pushes that tile's
// colour rows into the colour buffers of the other GPUs (TargetDesc::peer_color, 128-bit P2P stores), so the pass ends
// with the whole frame on every GPU and there is no separate gather: only tiles that were actually drawn cross NVLink,
// and the transfer overlaps the rasterisation of the other tiles. (Replicating every single colour store instead was
// measured first and is slower: 4-byte remote stores waste NVLink packets, and overdraw multiplies them.)
// Two cross-GPU barriers order the frame:
//   clear (each GPU clears ALL rows of its own colour buffer) -> barrier 1 -> k_raster (remote stores) -> barrier 2
// barrier 1: no GPU may store into a peer before that peer has cleared; barrier 2: the frame may only be read (or
// cleared again) once every peer's stores have landed. Each GPU owns an array of flag slots, one per peer, that the
// peers write through NVLink; epochs only grow, so re-running a barrier of an already passed epoch (replay of a pass
// after an arena overflow) neither blocks nor disturbs the others.
#pragma once
#include "rf_device.cuh"

#define RF_PEER_FLAG_STRIDE 32u  // uint32 words between flag slots (128 B)

struct PeerBarrier {
  uint32_t* flags[RF_MAX_PEERS + 1];  // flags[r]: rank r's slot array (device memory of GPU r); flags[self] is local
  uint32_t world, self;
};

__global__ void __launch_bounds__(32) k_peer_barrier(PeerBarrier B, uint32_t epoch) {
  const uint32_t r = threadIdx.x;
  if (r >= B.world || r == B.self) return;
  // stores of earlier kernels on this stream (the clear / the rasteriser's remote stores) are complete at the kernel
  // boundary; the system-scope fence orders them before the flag for observers on other GPUs
  __threadfence_system();
  uint32_t* theirs = B.flags[r] + B.self * RF_PEER_FLAG_STRIDE;
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
  const uint32_t* mine = B.flags[B.self] + r * RF_PEER_FLAG_STRIDE;
  uint32_t v;
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
  } while ((int32_t)(v - epoch) < 0);
}
