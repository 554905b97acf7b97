// tcgen05 (5th-generation tensor core) building blocks for sm_100a, inline PTX.
//
// FP32-accurate GEMM on the TF32 tensor pipe by operand splitting ("3xTF32"):
//   x = hi + lo,  hi = x with the low 13 mantissa bits cleared (exactly a TF32 value),
//                 lo = x - hi (exact in FP32; the MMA reads its top 10 mantissa bits)
//   A B ~= A_hi B_hi + A_lo B_hi + A_hi B_lo        (error ~2^-21 per product, FP32 accumulate)
//
// Operands live in shared memory in the canonical K-major, no-swizzle UMMA layout: 8-row x
// 16-byte "core matrices" (4 TF32 values per row); for a [rows][K] tile
//   byte(r, k) = (r / 8) * SBO + (k / 4) * LBO + (r % 8) * 16 + (k % 4) * 4,  LBO = 128, SBO = K * 32
// Accumulators live in tensor memory (TMEM, 512 columns x 128 lanes x 32 bit per SM); for
// M = 64 row i sits on lane 32 * (i / 16) + i % 16, for M = 128 on lane i.
// One elected thread issues tcgen05.mma; completion is signalled through tcgen05.commit on an
// mbarrier; epilogue threads read their own lane with tcgen05.ld.32x32b.
#pragma once
#include "macjd_common.cuh"

#ifndef MACJD_TEST_HOST_EMULATION
namespace macjd {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- operand split
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// ---- canonical K-major no-swizzle layout
__device__ __forceinline__ uint32_t umma_off_bytes(int r, int k, int K) {
  return (uint32_t)((r >> 3) * (K * 32) + (k >> 2) * 128 + (r & 7) * 16 + (k & 3) * 4);
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, leading /
// stride byte offsets (all >> 4), version 1 (Blackwell), layout type 0 (no swizzle)
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor) for kind::tf32, FP32 accumulate,
// A and B K-major
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// true in exactly one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---- TMEM management (one warp allocates and later frees)
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- fences / barriers
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }

// all previously issued tcgen05.mma of this thread arrive on `bar` when they complete
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
               : "memory");
}

// Bounded wait: a wrong descriptor or a peer that never arrives must neither hang the GPU nor kill the
// context: the thread raises the watchdog word (macjd_common.cuh) and carries on; the next library call
// reports MACJD_ERR_CUDA.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
#pragma unroll 1   // (nvcc unrolls this spin loop 32x otherwise: a third of the pair kernel's code)
  for (uint32_t it = 0; it < (1u << 22); ++it) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
  }
  watchdog_raise();
}

// ---- TMEM -> registers: 8 consecutive columns of this thread's lane
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
  v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
  v[4] = __uint_as_float(r4); v[5] = __uint_as_float(r5); v[6] = __uint_as_float(r6); v[7] = __uint_as_float(r7);
}

__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// 16 consecutive columns, no wait (pair with tmem_ld_wait)
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
// After tmem_ld_wait(): pins the uses of asynchronously loaded registers behind the wait
// (volatile asm statements keep their order; this one redefines v).
template <int N>
__device__ __forceinline__ void reg_fence(float (&v)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) asm volatile("" : "+f"(v[i]));
}

// 16 lanes x 256 bit x2: 16 consecutive columns of the 16 lanes starting at the warp's lane base,
// spread over all 32 threads like an m16n8 accumulator fragment: for column group i (8 columns),
// thread t holds v[4i+0..1] = (lane t/4, columns 8i + 2 (t%4) + {0,1}) and
// v[4i+2..3] = (lane t/4 + 8, same columns).
__device__ __forceinline__ void tmem_ld_16x256b_x2_nowait(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// x4: 32 consecutive columns, 16 registers (column group i = v[4i..4i+3])
__device__ __forceinline__ void tmem_ld_16x256b_x4_nowait(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// fast transcendental forms (MUFU ex2 / rcp): |error| ~1e-6 absolute on sigmoid / tanh
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }

// ---------------------------------------------------------------------------------------
// Self-test: D[M][N] = A[M][K] B[N][K]^T with the 3xTF32 split (one CTA, 128 threads).
// Exercises every primitive above; tests/test_gpu_tc05.py checks it against float64.
struct TcTestArgs {
  const float* A; const float* B; float* D;
  int M, N, K;
  int ld_shape;    // 0: tcgen05.ld.32x32b (thread = lane), 1: tcgen05.ld.16x256b (fragment layout, M = 64)
};

__global__ void __launch_bounds__(128, 1) tc_gemm_selftest_kernel(const TcTestArgs a) {
  extern __shared__ __align__(128) unsigned char tc_smem[];
  const int M = a.M, N = a.N, K = a.K;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* Ahi = reinterpret_cast<float*>(tc_smem);
  float* Alo = Ahi + (size_t)M * K;
  float* Bhi = Alo + (size_t)M * K;
  float* Blo = Bhi + (size_t)N * K;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;

  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  for (int idx = tid; idx < M * K; idx += 128) {
    const int r = idx / K, k = idx - r * K;
    const float x = a.A[idx], h = tf32_hi(x);
    const uint32_t off = umma_off_bytes(r, k, K) >> 2;
    Ahi[off] = h; Alo[off] = x - h;
  }
  for (int idx = tid; idx < N * K; idx += 128) {
    const int r = idx / K, k = idx - r * K;
    const float x = a.B[idx], h = tf32_hi(x);
    const uint32_t off = umma_off_bytes(r, k, K) >> 2;
    Bhi[off] = h; Blo[off] = x - h;
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;

  if (tid == 0) {
    const uint32_t idesc = umma_idesc_tf32(M, N);
    const uint32_t sbo = (uint32_t)K * 32, lbo = 128;
    const uint64_t dAhi = umma_smem_desc(smem_u32(Ahi), lbo, sbo), dAlo = umma_smem_desc(smem_u32(Alo), lbo, sbo);
    const uint64_t dBhi = umma_smem_desc(smem_u32(Bhi), lbo, sbo), dBlo = umma_smem_desc(smem_u32(Blo), lbo, sbo);
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint64_t adv = (uint64_t)((ks * 256) >> 4);      // 8 k = two 128-byte core matrices
      mma_tf32_ss(tmem_base, dAhi + adv, dBhi + adv, idesc, ks > 0);
      mma_tf32_ss(tmem_base, dAlo + adv, dBhi + adv, idesc, 1);
      mma_tf32_ss(tmem_base, dAhi + adv, dBlo + adv, idesc, 1);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();

  if (a.ld_shape == 1) {
    // fragment-layout read of the 16 valid lanes of this warp's quarter (M = 64 accumulators)
    for (int c0 = 0; c0 < N; c0 += 16) {
      float v[8];
      tmem_ld_16x256b_x2_nowait(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      reg_fence(v);
      for (int i = 0; i < 2; ++i)
        for (int hr = 0; hr < 2; ++hr)
          for (int j = 0; j < 2; ++j)
            a.D[(size_t)(warp * 16 + (lane >> 2) + 8 * hr) * N + c0 + 8 * i + 2 * (lane & 3) + j] = v[4 * i + 2 * hr + j];
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
    return;
  }
  // epilogue: warp w owns TMEM lanes 32w..32w+31
  const int row = (M == 128) ? tid : (warp * 16 + lane);
  const bool has_row = (M == 128) || (lane < 16);
  for (int c0 = 0; c0 < N; c0 += 8) {
    float v[8];
    tmem_ld8(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    if (has_row)
      for (int j = 0; j < 8; ++j) a.D[(size_t)row * N + c0 + j] = v[j];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

inline int tc_gemm_selftest(const macjd_ctx* ctx, int M, int N, int K, const float* A, const float* B, float* D,
                            int ld_shape = 0) {
  if (!ctx || !A || !B || !D) return MACJD_ERR_INVALID_ARG;
  if ((M != 64 && M != 128) || N < 16 || N > 256 || N % 16 || K < 8 || K % 8) return MACJD_ERR_UNSUPPORTED;
  const size_t smem = (size_t)2 * (M + N) * K * sizeof(float);
  if (smem > 200 * 1024) return MACJD_ERR_UNSUPPORTED;
  if (cudaFuncSetAttribute(tc_gemm_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return MACJD_ERR_CUDA;
  if (ld_shape == 1 && M != 64) return MACJD_ERR_UNSUPPORTED;
  TcTestArgs a{A, B, D, M, N, K, ld_shape};
  tc_gemm_selftest_kernel<<<1, 128, smem, (cudaStream_t)ctx->stream>>>(a);
  return MACJD_OK;
}

#if defined(MACJD_TC_PROFILE) || defined(MACJD_DEBUG_TOOLS)   // tooling build only (tools/tc_mma_rate.py)
// ---------------------------------------------------------------------------------------
// Micro-benchmark: issue `n` back-to-back tcgen05.mma (kind::tf32, SS) of shape M x N x 8 on
// garbage operands; out[0] = cycles to issue them, out[1] = cycles until the commit fires.
__global__ void __launch_bounds__(128, 1) tc_mma_rate_kernel(int M, int N, int n, unsigned long long* out) {
  extern __shared__ __align__(128) unsigned char tc_smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  for (int i = tid; i < 40 * 1024 / 4; i += 128) reinterpret_cast<float*>(tc_smem)[i] = 0.001f * (float)(i & 255);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  if (warp == 0) {
    // whole warp converged, one elected lane issues: operands stay in uniform registers (issued from
    // inside `if (tid == 0)` every MMA is wrapped in an ELECT / R2UR sequence that costs ~50 cycles)
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t idesc = umma_idesc_tf32(M, N);
    const uint64_t da = umma_smem_desc(smem_u32(tc_smem), 128, 16 * 32);
    const uint64_t db = umma_smem_desc(smem_u32(tc_smem) + 16384, 128, 16 * 32);
    const unsigned long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
      const uint32_t d = tb + (uint32_t)((i & 1) * 256);
      const uint64_t a = da + (uint64_t)((i & 1) * 16);
      if (elect_one()) mma_tf32_ss(d, a, db, idesc, 1u);
      __syncwarp();
    }
    if (elect_one()) mma_commit(&bar);
    __syncwarp();
    const unsigned long long t1 = clock64();
    mbar_wait(&bar, 0);
    const unsigned long long t2 = clock64();
    if (tid == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// Same with `issuers` (1..4) threads of different warps issuing n MMAs each into separate accumulators:
// out[0] = cycles until every issuer's commit has fired (tells whether the ~50-cycle floor per MMA
// belongs to the issuing thread or to the tensor pipe).
__global__ void __launch_bounds__(128, 1) tc_mma_rate_multi_kernel(int M, int N, int n, int issuers, unsigned long long* out) {
  extern __shared__ __align__(128) unsigned char tc_smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
  for (int i = tid; i < 40 * 1024 / 4; i += 128) reinterpret_cast<float*>(tc_smem)[i] = 0.001f * (float)(i & 255);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  const unsigned long long t0 = clock64();
  if ((tid & 31) == 0 && warp < issuers) {
    const uint32_t idesc = umma_idesc_tf32(M, N);
    const uint64_t da = umma_smem_desc(smem_u32(tc_smem), 128, 16 * 32);
    const uint64_t db = umma_smem_desc(smem_u32(tc_smem) + 16384, 128, 16 * 32);
    for (int i = 0; i < n; ++i) mma_tf32_ss(tmem_base + (uint32_t)(warp * 128), da + (uint64_t)((i & 1) * 16), db, idesc, 1u);
    mma_commit(&bar[warp]);
  }
  if (tid == 0) {
    for (int i = 0; i < issuers; ++i) mbar_wait(&bar[i], 0);
    out[0] = clock64() - t0;
    out[1] = out[0];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

inline int tc_mma_rate_multi(const macjd_ctx* ctx, int M, int N, int n, int issuers, unsigned long long* out_dev) {
  if (issuers < 1 || issuers > 4 || N > 128) return MACJD_ERR_INVALID_ARG;
  if (cudaFuncSetAttribute(tc_mma_rate_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024) != cudaSuccess)
    return MACJD_ERR_CUDA;
  tc_mma_rate_multi_kernel<<<1, 128, 64 * 1024, (cudaStream_t)ctx->stream>>>(M, N, n, issuers, out_dev);
  return MACJD_OK;
}

inline int tc_mma_rate(const macjd_ctx* ctx, int M, int N, int n, unsigned long long* out_dev) {
  if (cudaFuncSetAttribute(tc_mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024) != cudaSuccess)
    return MACJD_ERR_CUDA;
  tc_mma_rate_kernel<<<1, 128, 64 * 1024, (cudaStream_t)ctx->stream>>>(M, N, n, out_dev);
  return MACJD_OK;
}

#endif  // tooling build

}  // namespace tc
}  // namespace macjd
#endif  // !MACJD_TEST_HOST_EMULATION
