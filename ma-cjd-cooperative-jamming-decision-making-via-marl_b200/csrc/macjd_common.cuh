// Shared device helpers for the macjd kernels (sm_100a).
#pragma once

#ifdef MACJD_TEST_HOST_EMULATION
// CPU test suite only: the same sources compiled by g++ against tests/emul/cuda_emul.h.
#include "cuda_emul.h"
#else
#include <cuda_runtime.h>
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

__device__ __forceinline__ float sigmoidf_ref(float x) { return 1.0f / (1.0f + expf(-x)); }

}  // namespace macjd
