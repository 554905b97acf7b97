// Shared device helpers for the macjd kernels (sm_100a).
#pragma once

#ifdef MACJD_TEST_HOST_EMULATION
// CPU test suite only: the same sources compiled by g++ against tests/emul/cuda_emul.h.
#include "cuda_emul.h"
#else
#include <cuda_runtime.h>
#include <atomic>
// Every kernel launched through MACJD_LAUNCH starts with grid_dependency_wait() (below), so it may be
// launched as a programmatic dependent of its predecessor in the stream: its blocks are scheduled while
// the predecessor drains, the launch latency of the learner's ~60 short kernels overlaps, and stream
// order is still what the kernel observes.
namespace macjd {
template <typename... KArgs, typename... Args>
inline void launch_dependent(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
}  // namespace macjd
#define MACJD_LAUNCH(kernel, grid, block, smem, stream, ...) \
  ::macjd::launch_dependent(kernel, dim3(grid), dim3(block), (smem), (stream), __VA_ARGS__)
#define MACJD_DYNAMIC_SMEM(type, name)                             \
  extern __shared__ __align__(16) unsigned char macjd_dyn_smem_[]; \
  type* name = reinterpret_cast<type*>(macjd_dyn_smem_)
#endif

#include <stdint.h>

#include "../../include/macjd.h"

namespace macjd {

constexpr int kWarp = 32;
constexpr int kNumSMs = 148;  // B200

// ---------------------------------------------------------------- Philox4x32-10
// Counter-based RNG (Salmon et al. 2011).  Counter = (slot/4, step, index, stream),
// key = 64-bit seed.  The oracle reproduces the same mapping (oracle/env_oracle.py).
struct Philox4 { uint32_t x, y, z, w; };

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}

// uint32 -> uniform float in [0,1): top 24 bits, exact in fp32.
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

enum PhiloxStream : uint32_t { kStreamEnvNoise = 1, kStreamEpsilon = 2, kStreamRandomAction = 3, kStreamReplay = 4 };

__device__ __forceinline__ float philox_uniform(uint64_t seed, uint32_t stream, uint32_t index, uint32_t step,
                                                uint32_t slot) {
  const Philox4 r = philox4x32_10(slot >> 2, step, index, stream, (uint32_t)seed, (uint32_t)(seed >> 32));
  const uint32_t q = slot & 3u;
  return u01(q == 0 ? r.x : q == 1 ? r.y : q == 2 ? r.z : r.w);
}

// Fire-and-forget request that pulls a line into L2 (no register, no scoreboard entry).
__device__ __forceinline__ void prefetch_l2(const void* p) {
#ifndef MACJD_TEST_HOST_EMULATION
  asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p));
#else
  (void)p;
#endif
}

// Programmatic dependent launch (stream order kept, launch latency and independent prologue work
// overlapped): a kernel launched with the programmatic-serialization attribute may start while its
// predecessor still runs; it must call grid_dependency_wait() before touching anything the
// predecessor reads or writes.  A predecessor opts in with grid_launch_dependents(); without that
// (or without the launch attribute) both calls are no-ops and the launch is an ordinary one.
__device__ __forceinline__ void grid_dependency_wait() {
#ifndef MACJD_TEST_HOST_EMULATION
  asm volatile("griddepcontrol.wait;\n" ::: "memory");
#endif
}
__device__ __forceinline__ void grid_launch_dependents() {
#ifndef MACJD_TEST_HOST_EMULATION
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
#endif
}

// The learner's short kernels: let the NEXT kernel of the stream be launched right away -- its blocks are scheduled as
// this grid's blocks retire and park in their own grid_dependency_wait() until this grid has completed and flushed --
// then wait for the predecessor.  A train step at the reference batch is ~75 dependent launches of 3-20 us each; without
// the early trigger a dependent is only launched when the last block of its predecessor exits (~2 us per hand-over).
__device__ __forceinline__ void grid_dependency_sync() {
  grid_launch_dependents();
  grid_dependency_wait();
}

__device__ __forceinline__ float sigmoidf_ref(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---------------------------------------------------------------- pipeline watchdog
// The tensor-core kernels wait on mbarriers with a bound.  A wait that runs out (a peer CTA
// that never arrives: preemption, a debugger, a bug) must not kill the CUDA context -- the
// library's contract is status codes -- so instead of trapping the thread raises one word of
// page-locked host memory and carries on (every later wait is bounded too, so the kernel
// ends; its outputs are garbage).  The next entry point sees the word and returns
// MACJD_ERR_CUDA (macjd_api.cu: DeviceScope).  One word per process, created on first use.
#ifndef MACJD_TEST_HOST_EMULATION
inline int* watchdog_host_word() {
  static int* word = [] {
    int* p = nullptr;
    if (cudaHostAlloc(reinterpret_cast<void**>(&p), sizeof(int), cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
      cudaGetLastError();
      return static_cast<int*>(nullptr);
    }
    *p = 0;
    return p;
  }();                       // (thread-safe: C++11 static initialisation)
  return word;
}
inline int* g_watchdog_seen = nullptr;       // set once some launch has asked for the word (no allocation on plain calls)
inline bool watchdog_tripped() {
  int* w = g_watchdog_seen;
  return w && *reinterpret_cast<volatile int*>(w) != 0;
}
inline int* watchdog_device_word() {       // page-locked + unified addressing: the host pointer is the device pointer
  int* w = watchdog_host_word();
  g_watchdog_seen = w;
  return w;
}
// set on a device the first time one of its tensor-core kernels is launched (watchdog_arm)
__device__ int* g_watchdog_dev = nullptr;
__device__ __forceinline__ void watchdog_raise() {
  int* word = g_watchdog_dev;
  if (word) { *reinterpret_cast<volatile int*>(word) = 1; __threadfence_system(); }
}
#else
inline bool watchdog_tripped() { return false; }
#endif

// ---------------------------------------------------------------- per-device launch caches
// "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) was already requested for >= n bytes on device d":
// the attribute is sticky per device and asking costs microseconds on a latency-critical call.  Lock-free;
// two threads racing on the first call both make the (idempotent) request.
#ifndef MACJD_TEST_HOST_EMULATION
constexpr int kMaxCachedDevices = 64;
struct PerDeviceMax {
  std::atomic<size_t> v[kMaxCachedDevices];
  bool covers(int dev, size_t need) const {
    return dev >= 0 && dev < kMaxCachedDevices && v[dev].load(std::memory_order_relaxed) >= need;
  }
  void record(int dev, size_t need) {
    if (dev < 0 || dev >= kMaxCachedDevices) return;       // (uncached devices ask every time)
    size_t cur = v[dev].load(std::memory_order_relaxed);
    while (cur < need && !v[dev].compare_exchange_weak(cur, need, std::memory_order_relaxed)) {}
  }
};
// the pipeline watchdog word is published to a device once (a synchronous 8-byte symbol copy)
inline int watchdog_arm(int dev) {
  static PerDeviceMax armed;
  if (armed.covers(dev, 1)) return MACJD_OK;
  int* w = watchdog_device_word();
  if (w && cudaMemcpyToSymbol(g_watchdog_dev, &w, sizeof(w)) != cudaSuccess) return MACJD_ERR_CUDA;
  armed.record(dev, 1);
  return MACJD_OK;
}
#endif

}  // namespace macjd
