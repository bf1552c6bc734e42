// TEST INFRASTRUCTURE ONLY — a stand-in for <cuda_runtime.h> that lets g++ compile retrofire_b200/csrc/*.cu(h) for the
// host, so that the -m "not gpu" suite can execute the *kernel source* (its warp-level logic: ballots, shuffles,
// match_any dependency rounds, warp-aggregated allocation, the two-level sort ...) and compare it with the oracle in a
// container that has no GPU. It is NOT a product path: nothing under retrofire_b200/ can load it (the package only ever
// dlopens librf_b200.so, nvcc-built SASS), it is built by tests/emu/build_emu.py into tests/emu/_build/ and loaded by
// tests/test_emu_kernels.py alone. Speed is irrelevant here (a frame takes seconds); fidelity of the SIMT semantics is
// the point.
//
// Execution model: blocks run one after the other; the threads of a block are fibers on one OS thread. A fiber runs
// until it reaches a warp collective (__shfl_sync, __ballot_sync, __match_any_sync, __syncwarp ...) or __syncthreads(),
// where it waits — by switching to the next fiber — until every lane named in the mask has arrived. Between
// collectives the lanes of a warp are NOT in lock step (any interleaving the CUDA model allows for divergent
// threads), so code that relies on implicit warp synchrony fails here as it may on hardware. __activemask() returns
// only the calling lane (a legal outcome on Volta+ hardware). Arithmetic is the host's IEEE binary32 (build with
// -ffp-contract=off): + - * / sqrt are correctly rounded on both sides; powf comes from libm and may differ from
// CUDA's by an ulp, which is why the two powf shaders carry a +-1 LSB colour tolerance everywhere.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <type_traits>
#include <vector>

#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

#define RF_EMU 1
#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
// __noinline__ is rewritten by build_emu.py (glibc spells __attribute__((__noinline__)) itself, so it cannot be a macro)
#define __launch_bounds__(...)

// ---- vector types ------------------------------------------------------------------------------------------------
struct uint2 { uint32_t x, y; };
struct uint3 { uint32_t x, y, z; };
struct uint4 { uint32_t x, y, z, w; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

inline uint3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

template <class A, class B, class = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>
inline std::common_type_t<A, B> min(A a, B b) { using C = std::common_type_t<A, B>; return (C)a < (C)b ? (C)a : (C)b; }
template <class A, class B, class = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>
inline std::common_type_t<A, B> max(A a, B b) { using C = std::common_type_t<A, B>; return (C)a > (C)b ? (C)a : (C)b; }
// for arguments that only convert to a number (PaddedCounter)
inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }

// ---- the fiber scheduler -------------------------------------------------------------------------------------------
namespace emu {

#if defined(__x86_64__)
extern "C" void rf_emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.hidden rf_emu_switch
.globl rf_emu_switch
.type rf_emu_switch,@function
rf_emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size rf_emu_switch,.-rf_emu_switch
)");
#else
#error "tests/emu: the fiber switch is written for x86-64 only"
#endif

constexpr size_t kStackBytes = 256u << 10;
constexpr unsigned kMaxThreads = 1024;

struct Fiber {
  void* sp = nullptr;
  bool done = true;
};
struct GroupBarrier { uint32_t mask, arrived, gen; };
struct WarpState {
  uint64_t buf[32];
  std::vector<GroupBarrier> bars;  // one barrier per distinct mask: groups may overlap in time (a lane in a divergent
                                   // branch runs a sub-group collective while the others already wait at a full-warp one)
  uint32_t exited;
  GroupBarrier& bar(uint32_t mask) {
    for (GroupBarrier& b : bars) if (b.mask == mask) return b;
    bars.push_back(GroupBarrier{mask, 0u, 0u});
    return bars.back();
  }
};
struct BlockState {
  unsigned n = 0;
  unsigned arrived = 0, gen = 0, exited = 0;
  std::vector<WarpState> warps;
};

inline Fiber g_fibers[kMaxThreads];
inline unsigned char* g_stacks = nullptr;
inline void* g_sched_sp = nullptr;
inline unsigned g_cur = 0;
inline BlockState g_block;
inline unsigned long long g_progress = 0;  // bumped whenever a barrier completes or a fiber ends (deadlock detection)
inline void (*g_body)(void*) = nullptr;
inline void* g_body_arg = nullptr;
inline unsigned char* g_dyn_smem = nullptr;
inline const char* g_kernel = "?";
inline unsigned long long g_launches = 0;

inline void (*g_describe)(void*) = nullptr;  // set below: names the allocation next to a faulting address

inline void* dyn_smem() { return g_dyn_smem; }

inline void yield() { rf_emu_switch(&g_fibers[g_cur].sp, g_sched_sp); }

[[noreturn]] inline void die(const char* what) {
  std::fprintf(stderr, "rf emu: %s in kernel %s (block %u,%u thread %u)\n", what, g_kernel, blockIdx.x, blockIdx.y, g_cur);
  std::abort();
}

inline void fiber_main() {
  g_body(g_body_arg);
  const unsigned t = g_cur;
  g_fibers[t].done = true;
  g_block.warps[t >> 5].exited |= 1u << (t & 31);
  g_block.exited++;
  g_progress++;
  for (;;) yield();  // never returns: the scheduler does not switch to a finished fiber again
}

template <class F>
inline void run_block(unsigned nthreads, F& body) {
  if (nthreads == 0 || nthreads > kMaxThreads) die("bad block size");
  if (!g_stacks) g_stacks = static_cast<unsigned char*>(std::aligned_alloc(4096, kStackBytes * kMaxThreads));
  g_body = [](void* p) { (*static_cast<F*>(p))(); };
  g_body_arg = &body;
  g_block.n = nthreads;
  g_block.arrived = g_block.gen = g_block.exited = 0;
  g_block.warps.assign((nthreads + 31) / 32, WarpState{});
  for (unsigned w = 0; w < g_block.warps.size(); w++) {
    const unsigned live = std::min(32u, nthreads - w * 32);
    g_block.warps[w].exited = live == 32 ? 0u : ~((1u << live) - 1u);  // lanes that do not exist count as exited
  }
  for (unsigned t = 0; t < nthreads; t++) {
    uintptr_t top = reinterpret_cast<uintptr_t>(g_stacks + (size_t)(t + 1) * kStackBytes) & ~uintptr_t(15);
    void** sp = reinterpret_cast<void**>(top) - 8;  // r15 r14 r13 r12 rbx rbp ret pad
    for (int i = 0; i < 6; i++) sp[i] = nullptr;
    sp[6] = reinterpret_cast<void*>(&fiber_main);
    sp[7] = nullptr;
    g_fibers[t].sp = sp;
    g_fibers[t].done = false;
  }
  unsigned idle_rounds = 0;
  while (g_block.exited < nthreads) {
    const unsigned long long before = g_progress;
    for (unsigned t = 0; t < nthreads; t++) {
      if (g_fibers[t].done) continue;
      g_cur = t;
      threadIdx.x = t;
      rf_emu_switch(&g_sched_sp, g_fibers[t].sp);
    }
    if (g_progress == before) { if (++idle_rounds > 2) die("deadlock: every thread waits at a barrier that cannot complete"); }
    else idle_rounds = 0;
  }
}

// a wild access inside a kernel: say which kernel / block / thread, with a backtrace, before dying
inline void on_segv(int sig, siginfo_t* si, void*) {
  char msg[256];
  const int n = std::snprintf(msg, sizeof msg, "rf emu: signal %d at address %p in kernel %s (block %u,%u thread %u)\n", sig, si->si_addr, g_kernel,
                              blockIdx.x, blockIdx.y, g_cur);
  if (write(2, msg, (size_t)n) < 0) {}
  if (g_describe) g_describe(si->si_addr);
  void* bt[32];
  backtrace_symbols_fd(bt, backtrace(bt, 32), 2);
  _exit(139);
}
inline void install_handler() {
  static bool done = false;
  if (done) return;
  done = true;
  static unsigned char alt[64 << 10];
  stack_t ss{};
  ss.ss_sp = alt; ss.ss_size = sizeof alt;
  sigaltstack(&ss, nullptr);
  struct sigaction sa{};
  sa.sa_sigaction = on_segv;
  sa.sa_flags = SA_SIGINFO | SA_ONSTACK;
  sigaction(SIGSEGV, &sa, nullptr);
  sigaction(SIGBUS, &sa, nullptr);
}

template <class F>
inline void launch(const char* name, dim3 g, dim3 b, size_t smem, void* /*stream*/, F&& body) {
  install_handler();
  g_kernel = name;
  static const bool trace = std::getenv("RF_EMU_TRACE") != nullptr;
  if (trace) std::fprintf(stderr, "rf emu: launch %s grid %u,%u block %u smem %zu\n", g_kernel, g.x, g.y, b.x, smem);
  if (b.y != 1 || b.z != 1 || g.z != 1) die("only 1-D blocks and 2-D grids are emulated");
  std::vector<unsigned char> dyn(smem + 64, 0xA5);
  g_dyn_smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dyn.data()) + 63) & ~uintptr_t(63));
  gridDim = g; blockDim = b;
  g_launches++;
  for (unsigned by = 0; by < g.y; by++)
    for (unsigned bx = 0; bx < g.x; bx++) {
      blockIdx = uint3{bx, by, 0};
      run_block(b.x, body);
    }
  g_dyn_smem = nullptr;
}
template <class F> inline void launch(const char* name, dim3 g, dim3 b, size_t smem, F&& body) { launch(name, g, b, smem, nullptr, body); }
template <class F> inline void launch(const char* name, dim3 g, dim3 b, F&& body) { launch(name, g, b, 0, nullptr, body); }

// ---- barriers --------------------------------------------------------------------------------------------------------
inline void warp_barrier(uint32_t mask) {
  const unsigned t = g_cur, lane = t & 31;
  WarpState& w = g_block.warps[t >> 5];
  if (!(mask >> lane & 1u)) die("a lane calls a *_sync primitive with a mask that does not name it");
  if (mask == 1u << lane) return;
  const uint32_t gen = w.bar(mask).gen;
  w.bar(mask).arrived |= 1u << lane;
  for (;;) {
    GroupBarrier& b = w.bar(mask);  // looked up again after every yield: the vector may have grown
    if (b.gen != gen) return;
    if (((b.arrived | w.exited) & mask) == mask) { b.arrived = 0; b.gen++; g_progress++; return; }
    yield();
  }
}
inline void block_barrier() {
  const unsigned gen = g_block.gen;
  g_block.arrived++;
  for (;;) {
    if (g_block.gen != gen) return;
    if (g_block.arrived + g_block.exited >= g_block.n) { g_block.arrived = 0; g_block.gen++; g_progress++; return; }
    yield();
  }
}
template <class T> inline uint64_t to_bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, "shuffle payload"); std::memcpy(&b, &v, sizeof v); return b; }
template <class T> inline T from_bits(uint64_t b) { T v; std::memcpy(&v, &b, sizeof v); return v; }
// publish v, wait for the group, let `read` look at the group's values, wait again so nobody overwrites them early
template <class T, class R>
inline auto exchange(uint32_t mask, T v, R read) {
  WarpState& w = g_block.warps[g_cur >> 5];
  w.buf[g_cur & 31] = to_bits(v);
  warp_barrier(mask);
  auto r = read(w);
  warp_barrier(mask);
  return r;
}

}  // namespace emu

// ---- warp and block primitives -----------------------------------------------------------------------------------------
inline void __syncthreads() { emu::block_barrier(); }
inline void __syncwarp(uint32_t mask = 0xFFFFFFFFu) { emu::warp_barrier(mask); }
inline uint32_t __activemask() { return 1u << (emu::g_cur & 31); }
inline void __threadfence() {}
inline void __threadfence_block() {}
inline void __threadfence_system() {}

template <class T> inline T __shfl_sync(uint32_t mask, T v, int src, int width = 32) {
  const unsigned lane = emu::g_cur & 31;
  const unsigned s = (lane & ~(unsigned)(width - 1)) | ((unsigned)src & (unsigned)(width - 1));
  return emu::exchange(mask, v, [&](emu::WarpState& w) { return (mask >> s & 1u) ? emu::from_bits<T>(w.buf[s]) : v; });
}
template <class T> inline T __shfl_up_sync(uint32_t mask, T v, unsigned delta, int width = 32) {
  const unsigned lane = emu::g_cur & 31;
  const int s = (int)lane - (int)delta;
  const bool ok = s >= (int)(lane & ~(unsigned)(width - 1));
  return emu::exchange(mask, v, [&](emu::WarpState& w) { return ok && (mask >> s & 1u) ? emu::from_bits<T>(w.buf[s]) : v; });
}
template <class T> inline T __shfl_down_sync(uint32_t mask, T v, unsigned delta, int width = 32) {
  const unsigned lane = emu::g_cur & 31;
  const unsigned s = lane + delta;
  const bool ok = s < ((lane & ~(unsigned)(width - 1)) + (unsigned)width);
  return emu::exchange(mask, v, [&](emu::WarpState& w) { return ok && (mask >> s & 1u) ? emu::from_bits<T>(w.buf[s]) : v; });
}
template <class T> inline T __shfl_xor_sync(uint32_t mask, T v, int lanemask, int width = 32) {
  const unsigned lane = emu::g_cur & 31;
  const unsigned s = lane ^ (unsigned)lanemask;
  const bool ok = (s & ~(unsigned)(width - 1)) == (lane & ~(unsigned)(width - 1));
  return emu::exchange(mask, v, [&](emu::WarpState& w) { return ok && (mask >> s & 1u) ? emu::from_bits<T>(w.buf[s]) : v; });
}
inline uint32_t __ballot_sync(uint32_t mask, int pred) {
  return emu::exchange(mask, (uint32_t)(pred != 0), [&](emu::WarpState& w) {
    uint32_t r = 0;
    for (unsigned l = 0; l < 32; l++) if ((mask >> l & 1u) && !(w.exited >> l & 1u) && w.buf[l]) r |= 1u << l;
    return r;
  });
}
inline int __all_sync(uint32_t mask, int pred) {
  return emu::exchange(mask, (uint32_t)(pred != 0), [&](emu::WarpState& w) {
    for (unsigned l = 0; l < 32; l++) if ((mask >> l & 1u) && !(w.exited >> l & 1u) && !w.buf[l]) return 0;
    return 1;
  });
}
inline int __any_sync(uint32_t mask, int pred) {
  return emu::exchange(mask, (uint32_t)(pred != 0), [&](emu::WarpState& w) {
    for (unsigned l = 0; l < 32; l++) if ((mask >> l & 1u) && !(w.exited >> l & 1u) && w.buf[l]) return 1;
    return 0;
  });
}
inline uint32_t __reduce_add_sync(uint32_t mask, uint32_t v) {  // REDUX.SUM
  return emu::exchange(mask, v, [&](emu::WarpState& w) {
    uint32_t r = 0;
    for (unsigned l = 0; l < 32; l++) if ((mask >> l & 1u) && !(w.exited >> l & 1u)) r += (uint32_t)w.buf[l];
    return r;
  });
}
inline uint32_t __reduce_max_sync(uint32_t mask, uint32_t v) {  // REDUX.MAX
  return emu::exchange(mask, v, [&](emu::WarpState& w) {
    uint32_t r = 0;
    for (unsigned l = 0; l < 32; l++) if ((mask >> l & 1u) && !(w.exited >> l & 1u)) r = std::max(r, (uint32_t)w.buf[l]);
    return r;
  });
}
template <class T> inline uint32_t __match_any_sync(uint32_t mask, T v) {
  const uint64_t mine = emu::to_bits(v);
  return emu::exchange(mask, v, [&](emu::WarpState& w) {
    uint32_t r = 0;
    for (unsigned l = 0; l < 32; l++) if ((mask >> l & 1u) && !(w.exited >> l & 1u) && w.buf[l] == mine) r |= 1u << l;
    return r;
  });
}

// ---- scalar intrinsics -----------------------------------------------------------------------------------------------
template <class T> inline T __ldg(const T* p) { return *p; }
inline uint32_t __float_as_uint(float f) { return emu::from_bits<uint32_t>(emu::to_bits(f)); }
inline int32_t __float_as_int(float f) { return emu::from_bits<int32_t>(emu::to_bits(f)); }
inline float __uint_as_float(uint32_t u) { return emu::from_bits<float>(emu::to_bits(u)); }
inline float __int_as_float(int32_t u) { return emu::from_bits<float>(emu::to_bits(u)); }
// PTX cvt.rzi: round towards zero, clamp to the destination range, NaN -> 0
inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t sel) {  // PRMT, default mode
  const uint64_t src = (uint64_t)b << 32 | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; i++) r |= (uint32_t)((src >> (8 * ((sel >> (4 * i)) & 7u))) & 0xFFu) << (8 * i);
  return r;
}
inline uint32_t __float2uint_rz(float f) { return !(f == f) ? 0u : (f <= 0.0f ? 0u : (f >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)f)); }
inline int32_t __float2int_rz(float f) { return !(f == f) ? 0 : (f <= -2147483648.0f ? INT32_MIN : (f >= 2147483648.0f ? INT32_MAX : (int32_t)f)); }
inline int __popc(uint32_t v) { return __builtin_popcount(v); }
inline int __ffs(uint32_t v) { return __builtin_ffs((int)v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(uint32_t v) { return v ? __builtin_clz(v) : 32; }

template <class T, class U> inline T atomicAdd(T* p, U v) { const T o = *p; *p = (T)(o + (T)v); return o; }
template <class T, class U> inline T atomicOr(T* p, U v) { const T o = *p; *p = (T)(o | (T)v); return o; }
template <class T, class U> inline T atomicMax(T* p, U v) { const T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <class T, class U> inline T atomicMin(T* p, U v) { const T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <class T, class U> inline T atomicExch(T* p, U v) { const T o = *p; *p = (T)v; return o; }

// ---- the runtime API, as far as rf_api.cu uses it ------------------------------------------------------------------------
enum cudaError_t { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1, cudaErrorNotSupported = 801, cudaErrorPeerAccessAlreadyEnabled = 704 };
typedef cudaError_t cudaError;
struct CUstream_st { int id; };
typedef CUstream_st* cudaStream_t;
struct CUevent_st { std::chrono::steady_clock::time_point t; };
typedef CUevent_st* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void* devicePointer; void* hostPointer; };
struct cudaDeviceProp { char name[256]; int major, minor, multiProcessorCount; };
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

namespace emu {
inline std::map<uintptr_t, std::pair<size_t, cudaMemoryType>>& allocs() { static std::map<uintptr_t, std::pair<size_t, cudaMemoryType>> m; return m; }
inline void describe(void* addr);
inline cudaError_t alloc(void** p, size_t n, cudaMemoryType kind) {
  g_describe = &describe;
  void* q = nullptr;
  if (posix_memalign(&q, 256, n ? n : 1) != 0) return cudaErrorMemoryAllocation;
  std::memset(q, 0xCD, n);  // device memory is not zeroed
  allocs()[reinterpret_cast<uintptr_t>(q)] = {n, kind};
  *p = q;
  return cudaSuccess;
}
inline void describe(void* addr) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(addr);
  auto it = allocs().upper_bound(a);
  if (it == allocs().begin()) return;
  --it;
  std::fprintf(stderr, "rf emu: nearest allocation below: base %p, %zu bytes (%s); the address is %lld bytes past its end\n", reinterpret_cast<void*>(it->first),
               it->second.first, it->second.second == cudaMemoryTypeDevice ? "device" : "pinned host", (long long)(a - it->first) - (long long)it->second.first);
}
inline cudaError_t release(void* p) { if (p) { allocs().erase(reinterpret_cast<uintptr_t>(p)); std::free(p); } return cudaSuccess; }
inline int sm_count() { const char* e = std::getenv("RF_EMU_SMS"); const int n = e ? std::atoi(e) : 2; return n > 0 ? n : 2; }
}  // namespace emu

inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { std::memset(p, 0, sizeof *p); std::strcpy(p->name, "rf-emu"); p->major = 10; p->minor = 0; p->multiProcessorCount = emu::sm_count(); return cudaSuccess; }
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { return emu::alloc(reinterpret_cast<void**>(p), n, cudaMemoryTypeDevice); }
inline cudaError_t cudaFree(void* p) { return emu::release(p); }
template <class T> inline cudaError_t cudaHostAlloc(T** p, size_t n, unsigned) { return emu::alloc(reinterpret_cast<void**>(p), n, cudaMemoryTypeHost); }
inline cudaError_t cudaFreeHost(void* p) { return emu::release(p); }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t = nullptr) {
  for (size_t y = 0; y < h; y++) std::memmove(static_cast<char*>(d) + y * dp, static_cast<const char*>(s) + y * sp, w);
  return cudaSuccess;
}
inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new CUstream_st{0}; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new CUevent_st{std::chrono::steady_clock::now()}; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaMemGetInfo(size_t* fr, size_t* tot) { *fr = *tot = 0; return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return cudaSuccess; }
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
  std::memset(a, 0, sizeof *a);
  auto& m = emu::allocs();
  auto it = m.upper_bound(reinterpret_cast<uintptr_t>(p));
  if (it != m.begin()) {
    --it;
    if (reinterpret_cast<uintptr_t>(p) < it->first + std::max<size_t>(it->second.first, 1)) a->type = it->second.second;
  }
  return cudaSuccess;
}
template <class K> inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*) { return cudaErrorNotSupported; }
inline cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
