// Host emulation of the small CUDA subset the macjd kernels use.  TEST INFRASTRUCTURE.
//
// There is no GPU in the build container and GPU time is scarce, so the CPU test
// suite compiles the *same* kernel sources (csrc/*.cuh, csrc/macjd_api.cu) with g++
// against this header and runs them on host memory: every CUDA thread of a block is a
// ucontext fiber, __syncthreads()/__shfl_*_sync() are cooperative barriers, blocks run
// one after the other.  It exists to catch indexing / logic errors before a kernel
// ever reaches the B200; it is never built into or loaded by the product package
// (the product loader only opens libmacjd_b200.so and raises if it is missing).
#pragma once
#ifndef MACJD_TEST_HOST_EMULATION
#error "cuda_emul.h is test-only; compile with -DMACJD_TEST_HOST_EMULATION"
#endif

#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __shared__ static
#define __constant__ static
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __grid_constant__

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_e { unsigned x, y, z; };

struct float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct uint2 { unsigned x, y; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct __attribute__((aligned(16))) double2 { double x, y; };
static inline float2 make_float2(float x, float y) { return {x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return {x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return {x, y, z, w}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return {x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return {x, y}; }

typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
typedef void* cudaEvent_t;
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { static int tag; *s = &tag; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { static int tag; *e = &tag; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }

namespace emul {

struct State {
  dim3 threadIdx, blockIdx, blockDim, gridDim;
  int nthreads = 0;
  int cur = 0;                 // fiber currently running
  std::vector<char> done;
  // block barrier
  int bar_count = 0;
  unsigned bar_gen = 0;
  // warp barriers / exchange slots
  std::vector<int> wbar_count;
  std::vector<unsigned> wbar_gen;
  std::vector<uint64_t> xchg;  // [nthreads]
  unsigned char* dyn_smem = nullptr;
  ucontext_t sched_ctx;
  std::vector<ucontext_t> ctx;
  std::vector<std::vector<unsigned char>> stacks;
  std::function<void()> body;
};
inline State& S() { static State s; return s; }

inline void yield_to_scheduler() {
  State& s = S();
  int me = s.cur;
  swapcontext(&s.ctx[me], &s.sched_ctx);
}

inline void block_barrier() {
  State& s = S();
  unsigned gen = s.bar_gen;
  if (++s.bar_count == s.nthreads) {
    s.bar_count = 0;
    s.bar_gen++;
    return;
  }
  while (s.bar_gen == gen) yield_to_scheduler();
}

inline int warp_width(int warp) {
  State& s = S();
  return std::min(32, s.nthreads - warp * 32);
}

inline void warp_barrier() {
  State& s = S();
  int me = s.cur, w = me / 32;
  unsigned gen = s.wbar_gen[w];
  if (++s.wbar_count[w] == warp_width(w)) {
    s.wbar_count[w] = 0;
    s.wbar_gen[w]++;
    return;
  }
  while (s.wbar_gen[w] == gen) yield_to_scheduler();
}

template <class T>
inline T shfl_from(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle of <= 8 byte types only");
  State& s = S();
  int me = s.cur, w = me / 32;
  uint64_t bits = 0;
  memcpy(&bits, &v, sizeof(T));
  s.xchg[me] = bits;
  warp_barrier();
  int width = warp_width(w);
  int src = (src_lane >= 0 && src_lane < width) ? w * 32 + src_lane : me;
  uint64_t got = s.xchg[src];
  warp_barrier();
  T out;
  memcpy(&out, &got, sizeof(T));
  return out;
}

static void trampoline() {
  State& s = S();
  s.body();
  s.done[s.cur] = 1;
  // return to the scheduler for good
  swapcontext(&s.ctx[s.cur], &s.sched_ctx);
}

template <class Kernel, class... Args>
void launch(Kernel kernel, dim3 grid, dim3 block, size_t smem_bytes, Args... args) {
  State& s = S();
  const int nt = (int)(block.x * block.y * block.z);
  const size_t kStack = 256 * 1024;
  s.nthreads = nt;
  s.blockDim = block;
  s.gridDim = grid;
  if ((int)s.stacks.size() < nt) s.stacks.resize(nt);
  for (int i = 0; i < nt; ++i)
    if (s.stacks[i].size() != kStack) s.stacks[i].resize(kStack);
  s.ctx.resize(nt);
  s.done.assign(nt, 0);
  s.xchg.assign(nt, 0);
  int nwarps = (nt + 31) / 32;
  std::vector<unsigned char> dyn(smem_bytes + 64);
  s.dyn_smem = (unsigned char*)(((uintptr_t)dyn.data() + 63) & ~(uintptr_t)63);
  s.body = [&]() { kernel(args...); };
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        s.blockIdx = dim3(bx, by, bz);
        s.bar_count = 0;
        s.wbar_count.assign(nwarps, 0);
        s.wbar_gen.assign(nwarps, 0);
        std::fill(s.done.begin(), s.done.end(), 0);
        for (int i = 0; i < nt; ++i) {
          getcontext(&s.ctx[i]);
          s.ctx[i].uc_stack.ss_sp = s.stacks[i].data();
          s.ctx[i].uc_stack.ss_size = kStack;
          s.ctx[i].uc_link = &s.sched_ctx;
          makecontext(&s.ctx[i], (void (*)())trampoline, 0);
        }
        int remaining = nt;
        while (remaining > 0) {
          int progressed = 0;
          for (int i = 0; i < nt; ++i) {
            if (s.done[i]) continue;
            s.cur = i;
            s.threadIdx = dim3(i % block.x, (i / block.x) % block.y, i / (block.x * block.y));
            swapcontext(&s.sched_ctx, &s.ctx[i]);
            progressed = 1;
            if (s.done[i]) --remaining;
          }
          if (!progressed) break;
        }
      }
  s.dyn_smem = nullptr;
}

}  // namespace emul

#define threadIdx (emul::S().threadIdx)
#define blockIdx (emul::S().blockIdx)
#define blockDim (emul::S().blockDim)
#define gridDim (emul::S().gridDim)

static inline void __syncthreads() { emul::block_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emul::warp_barrier(); }
template <class T> static inline T __shfl_sync(unsigned, T v, int lane, int = 32) { return emul::shfl_from(v, lane); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) {
  return emul::shfl_from(v, (emul::S().cur % 32) ^ m);
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d, int = 32) {
  return emul::shfl_from(v, (emul::S().cur % 32) + d);
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, int d, int = 32) {
  return emul::shfl_from(v, (emul::S().cur % 32) - d);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  unsigned mine = pred ? (1u << (emul::S().cur % 32)) : 0u, all = 0;
  for (int l = 0; l < 32; ++l) all |= emul::shfl_from(mine, l);
  return all;
}

template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcs(const T* p) { return *p; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }

static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
#define __expf(x) expf(x)
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
static inline float __saturatef(float a) { return fminf(fmaxf(a, 0.f), 1.f); }
static inline int __float2int_rn(float a) { return (int)lrintf(a); }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }

template <class T> static inline T atomicAdd(T* p, T v) { T old = *p; *p = old + v; return old; }
static inline int atomicMax(int* p, int v) { int old = *p; if (v > old) *p = v; return old; }
static inline unsigned atomicInc(unsigned* p, unsigned lim) { unsigned old = *p; *p = (old >= lim) ? 0 : old + 1; return old; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline void __threadfence() {}
static inline void __trap() { abort(); }

#define MACJD_LAUNCH(kernel, grid, block, smem, stream, ...) \
  emul::launch(kernel, dim3(grid), dim3(block), (size_t)(smem), __VA_ARGS__)
#define MACJD_DYNAMIC_SMEM(type, name) type* name = reinterpret_cast<type*>(emul::S().dyn_smem)
