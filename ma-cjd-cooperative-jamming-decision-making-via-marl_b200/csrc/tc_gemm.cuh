// General GEMM of the learner on the tcgen05 tensor cores (3xTF32: FP32-level accuracy), same contract as
// sgemm_kernel (sgemm.cuh: GemmArgs -- operand orientations ta / tb, bias, ReLU / clamp / sigmoid, mask, accumulate,
// deterministic split-K partials).  ncu on the FP32 SIMT GEMM at the stress shapes (BASELINE config 4: 101 376 mixer
// rows, 202 752 Q-head rows, K = 128 / 256) shows it FMA-issue bound at ~15 TFLOP/s -- dense contractions, which is
// where north_star sends the work to the tensor cores.
//
// One CTA = one 128 x 128 output tile (cta_group::1, M = 128, N = 128 MMAs, accumulator = 128 TMEM columns).
//   warps 0-7  stage the operands: global FP32 -> registers -> TF32 hi / lo split -> shared memory in the K-major
//              no-swizzle UMMA layout (tc05.cuh), 16 k per stage, three stages (32 KB each: two CTAs share an SM, one
//              loads while the other computes); the next chunk's global loads are in flight while the current one is
//              split and stored; either operand may be
//              k-contiguous in memory (16-byte loads along k) or row-contiguous (16-byte loads along the rows: the
//              transposed forms of the weight-gradient products, whose contraction runs over the batch rows);
//              afterwards they are the epilogue (thread = row; warps w and w + 4 split the 128 columns)
//   warp 8     issues the MMAs (whole warp converged, one elected lane: operands in uniform registers) and commits
//              each stage back to the loaders
#pragma once
#include <stdlib.h>
#include "sgemm.cuh"
#include "tc05.cuh"

#ifndef MACJD_TEST_HOST_EMULATION
namespace macjd {
namespace tc {

constexpr int kGtM = 128, kGtN = 128, kGtK = 16, kGtStages = 3;     // 96 KB of stages: two CTAs per SM
constexpr int kGtLoadThreads = 256, kGtThreads = kGtLoadThreads + 32;
constexpr int kGtTileFloats = kGtM * kGtK;                       // one hi (or lo) operand tile of a stage
constexpr int kGtTileLd = kGtN + 4;                              // epilogue staging tile: padded rows (bank spread, 16-byte aligned)

struct GtSmem {
  float a[kGtStages][2][kGtTileFloats];                          // [stage][hi / lo]
  float b[kGtStages][2][kGtTileFloats];
  uint64_t full[kGtStages], empty[kGtStages], done;
  uint32_t tmem_base;
};

// write 4 consecutive k of one row (hi and lo parts) into a [128][kGtK] operand tile
__device__ __forceinline__ void store_split4_tile(float* hi, float* lo, int r, int k, const float (&v)[4]) {
  const uint32_t off = umma_off_bytes(r, k, kGtK) >> 2;
  float4 h, l;
  h.x = tf32_hi(v[0]); h.y = tf32_hi(v[1]); h.z = tf32_hi(v[2]); h.w = tf32_hi(v[3]);
  l.x = v[0] - h.x; l.y = v[1] - h.y; l.z = v[2] - h.z; l.w = v[3] - h.w;
  *reinterpret_cast<float4*>(hi + off) = h;
  *reinterpret_cast<float4*>(lo + off) = l;
}

__device__ __forceinline__ void gt_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// One 128 x 16 operand tile, element (row, k) = src(row0 + row, k0 + k), zero outside the matrix, in two halves so
// that the global loads of the NEXT chunk are in flight while this chunk is split and stored:
//   gt_load : global -> registers (kGtVec 16-byte vectors per thread)
//   gt_store: registers -> TF32 hi / lo -> shared memory (UMMA K-major layout)
// kContigK: src(r, k) = p[r * ld + k] (vectors along k); else src(r, k) = p[k * ld + r] (vectors along the rows).
constexpr int kGtVec = kGtTileFloats / 4 / kGtLoadThreads;      // 2
constexpr int kGtAhead = 4;                                      // chunks in flight per thread

template <bool kContigK>
__device__ __forceinline__ void gt_load(float4 (&x)[kGtVec], const float* p, int ld, int row0, int n_rows, int k0, int kend, int tid,
                                        bool vec_ok) {
#pragma unroll
  for (int q = 0; q < kGtVec; ++q) {
    const int idx = tid + q * kGtLoadThreads;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (kContigK) {
      // a warp takes 8 rows x 16 k: lane -> (row % 8 = lane % 8, k / 4 = lane / 8), so that the 8 lanes of a quarter-warp
      // store 128 contiguous bytes of the tile (8 rows of one k-group); a row's four 16-byte pieces stay in one warp load
      const int r = ((idx >> 5) << 3) + (idx & 7), k = ((idx >> 3) & 3) << 2;
      const int gr = row0 + r, gk = k0 + k;
      if (gr < n_rows) {
        const float* s = p + (size_t)gr * ld + gk;
        if (vec_ok && gk + 3 < kend) {
          x[q] = __ldg(reinterpret_cast<const float4*>(s));
          continue;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (gk + j < kend) v[j] = __ldg(s + j);
      }
    } else {
      // four core matrices (8 rows x 4 k = 128 contiguous bytes of the tile each) per vector slot: lane -> (row % 8 =
      // lane / 4, k % 4 = lane % 4); the warp's 32 loads cover 4 source rows x 32 contiguous bytes
      const int warp_ = tid >> 5, lane_ = tid & 31;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int blk = (q * 4 + j) * (kGtLoadThreads / 32) + warp_;       // 0 .. 63 = (row block 0..15) x (k block 0..3)
        const int r = (blk >> 2) * 8 + (lane_ >> 2), k = (blk & 3) * 4 + (lane_ & 3);
        const int gr = row0 + r, gk = k0 + k;
        if (gr < n_rows && gk < kend) v[j] = __ldg(p + (size_t)gk * ld + gr);
      }
    }
    x[q] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

template <bool kContigK>
__device__ __forceinline__ void gt_store(float* hi, float* lo, const float4 (&x)[kGtVec], int tid) {
#pragma unroll
  for (int q = 0; q < kGtVec; ++q) {
    const int idx = tid + q * kGtLoadThreads;
    const float v[4] = {x[q].x, x[q].y, x[q].z, x[q].w};
    if (kContigK) {
      store_split4_tile(hi, lo, ((idx >> 5) << 3) + (idx & 7), ((idx >> 3) & 3) << 2, v);
    } else {
      const int warp_ = tid >> 5, lane_ = tid & 31;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int blk = (q * 4 + j) * (kGtLoadThreads / 32) + warp_;
        const float h = tf32_hi(v[j]);
        hi[blk * 32 + lane_] = h;                 // core matrix blk, element (lane / 4, lane % 4): conflict-free
        lo[blk * 32 + lane_] = v[j] - h;
      }
    }
  }
}

__device__ __forceinline__ void gt_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;\n" ::"r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void gt_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// B once per call instead of once per CTA.  In a tall product (the learner's layers over T x M rows: 1 600 row tiles) every
// CTA of a column used to load the same FP32 weights, split them to TF32 hi / lo and lay them out for the MMA -- half of
// the loader warps' work, and the loaders bound the kernel (ncu r2s: tensor pipe active 19 %, issue active 34 %).  This
// kernel writes, per 128-column tile and 16-k chunk, the hi tile followed by the lo tile exactly as a stage's B buffers
// hold them (16 KB, zero padded); the GEMM then fetches a stage's B with ONE bulk copy.
constexpr int kGtPackFloats = 2 * kGtTileFloats;                 // hi + lo of one chunk
inline size_t tc_pack_b_floats(int N, int K) { return (size_t)((N + kGtN - 1) / kGtN) * ((K + kGtK - 1) / kGtK) * kGtPackFloats; }

__global__ void __launch_bounds__(256) tc_pack_b_kernel(const float* __restrict__ B, int ldb, int tb, int N, int K, float* __restrict__ out) {
  grid_dependency_sync();               // (the previous GEMM of the stream may still be reading `out`)
  const int kchunks = (K + kGtK - 1) / kGtK;
  const int nt = blockIdx.x / kchunks, kc = blockIdx.x - nt * kchunks;
  const int r = threadIdx.x >> 1, kh = (threadIdx.x & 1) * 8;     // column of the tile, 8 of its 16 k
  const int n = nt * kGtN + r;
  float* hi = out + (size_t)blockIdx.x * kGtPackFloats;
  float* lo = hi + kGtTileFloats;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = kc * kGtK + kh + 4 * q + j;
      v[j] = (n < N && k < K) ? (tb ? __ldg(B + (size_t)n * ldb + k) : __ldg(B + (size_t)k * ldb + n)) : 0.f;
    }
    store_split4_tile(hi, lo, r, kh + 4 * q, v);
  }
}

template <bool kAContigK, bool kBContigK, bool kBPacked = false>
__global__ void __launch_bounds__(kGtThreads, 2) tc_gemm_kernel(const GemmArgs g) {
  grid_launch_dependents();             // (the next kernel of the stream may be launched; it waits for this grid before its first access)
  extern __shared__ __align__(1024) unsigned char gt_raw[];
  GtSmem& S = *reinterpret_cast<GtSmem*>(gt_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * kGtM, n0 = blockIdx.x * kGtN;
  const int kbeg = blockIdx.z * g.k_per_split;
  const int kend = min(g.K, kbeg + g.k_per_split);
  const int n_chunks = (kend - kbeg + kGtK - 1) / kGtK;
  if (tid == 0) {
    for (int s = 0; s < kGtStages; ++s) { mbar_init(&S.full[s], kGtLoadThreads + (kBPacked ? 1 : 0)); mbar_init(&S.empty[s], 1); }
    mbar_init(&S.done, 1);
    fence_mbar_init();
  }
  if (warp == kGtLoadThreads / 32) tmem_alloc(&S.tmem_base, kGtN);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = S.tmem_base;
  grid_dependency_wait();               // (operands may come from the preceding kernel of the stream)

  if (warp < kGtLoadThreads / 32) {
    // ================================================================== loaders, then epilogue
    // 16-byte loads need aligned rows: base pointer and leading dimension multiples of four floats
    const bool a_vec = ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0) && (g.lda & 3) == 0 && (!kAContigK || (kbeg & 3) == 0);
    const bool b_vec = ((reinterpret_cast<uintptr_t>(g.B) & 15) == 0) && (g.ldb & 3) == 0 && (!kBContigK || (kbeg & 3) == 0);
    uint32_t empty_par = 0;
    // kGtAhead chunks of both operands are in flight per thread (registers) while the oldest one is converted and
    // stored: at ~1.5 us of memory latency under load a single chunk in flight left every k-step waiting for DRAM / L2
    float4 ra[kGtAhead][kGtVec], rb[kBPacked ? 1 : kGtAhead][kGtVec];
    // kBPacked: a stage's B (hi + lo tiles, 16 KB, contiguous in S.b[s]) arrives by one bulk copy from the packed buffer
    const float* bsrc = kBPacked ? g.bpack + ((size_t)blockIdx.x * g.bpack_kchunks + kbeg / kGtK) * kGtPackFloats : nullptr;
#pragma unroll
    for (int d = 0; d < kGtAhead; ++d)
      if (d < n_chunks) {
        gt_load<kAContigK>(ra[d], g.A, g.lda, m0, g.M, kbeg + d * kGtK, kend, tid, a_vec);
        if (!kBPacked) gt_load<kBContigK>(rb[d < (kBPacked ? 1 : kGtAhead) ? d : 0], g.B, g.ldb, n0, g.N, kbeg + d * kGtK, kend, tid, b_vec);
      }
    for (int c0 = 0; c0 < n_chunks; c0 += kGtAhead) {
#pragma unroll
      for (int d = 0; d < kGtAhead; ++d) {
        const int c = c0 + d;
        if (c < n_chunks) {
          const int s = c % kGtStages;
          if (c >= kGtStages) { mbar_wait(&S.empty[s], (empty_par >> s) & 1u); empty_par ^= 1u << s; }
          if (kBPacked && tid == 0) {
            gt_expect_tx(&S.full[s], (uint32_t)(kGtPackFloats * 4));
            gt_bulk_g2s(S.b[s][0], bsrc + (size_t)c * kGtPackFloats, (uint32_t)(kGtPackFloats * 4), &S.full[s]);
          }
          gt_store<kAContigK>(S.a[s][0], S.a[s][1], ra[d], tid);
          if (!kBPacked) gt_store<kBContigK>(S.b[s][0], S.b[s][1], rb[d < (kBPacked ? 1 : kGtAhead) ? d : 0], tid);
          fence_async_smem();
          gt_mbar_arrive(&S.full[s]);
          if (c + kGtAhead < n_chunks) {           // refill the slot just consumed
            gt_load<kAContigK>(ra[d], g.A, g.lda, m0, g.M, kbeg + (c + kGtAhead) * kGtK, kend, tid, a_vec);
            if (!kBPacked) gt_load<kBContigK>(rb[d < (kBPacked ? 1 : kGtAhead) ? d : 0], g.B, g.ldb, n0, g.N, kbeg + (c + kGtAhead) * kGtK, kend, tid, b_vec);
          }
        }
      }
    }
    mbar_wait(&S.done, 0);
    fence_after_sync();
    // ---- epilogue.  The accumulator is read thread = row (a TMEM lane); written that way every store instruction
    // of a warp would touch 32 different rows, 4 bytes each.  So the tile goes through shared memory (the operand
    // stages are free now: every MMA has completed) and leaves in row order: a warp writes 512 contiguous bytes.
    float* tile = reinterpret_cast<float*>(gt_raw);                     // [128][kGtTileLd]
    {
      const int q4 = warp & 3, half = warp >> 2;
      const int r = q4 * 32 + lane;
      const uint32_t tl = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(half * 64);
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16) {
        float v[16];
        tmem_ld16_nowait(tl + (uint32_t)c0, v);
        tmem_ld_wait();
        reg_fence(v);
        float4* dst = reinterpret_cast<float4*>(tile + (size_t)r * kGtTileLd + half * 64 + c0);
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
    }
    fence_before_sync();
    asm volatile("bar.sync 1, %0;\n" ::"n"(kGtLoadThreads) : "memory");
    const bool part = gridDim.z > 1;
    float* out = part ? g.partial + (size_t)blockIdx.z * g.M * g.N : g.C;
    const int ldo = part ? g.N : g.ldc;
    const bool o_vec = ((reinterpret_cast<uintptr_t>(out) & 15) == 0) && (ldo & 3) == 0 &&
                       (!g.mask || (((reinterpret_cast<uintptr_t>(g.mask) & 15) == 0) && (g.ldmask & 3) == 0)) &&
                       (!g.bias || (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0);
#pragma unroll 2
    for (int idx = tid; idx < kGtM * (kGtN / 4); idx += kGtLoadThreads) {
      const int r = idx >> 5, c = (idx & 31) << 2;
      const int m = m0 + r, n = n0 + c;
      if (m >= g.M || n >= g.N) continue;
      const float4 t4 = *reinterpret_cast<const float4*>(tile + (size_t)r * kGtTileLd + c);
      float x[4] = {t4.x, t4.y, t4.z, t4.w};
      float* dst = out + (size_t)m * ldo + n;
      const bool full4 = o_vec && n + 3 < g.N;
      if (!part) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (n + j >= g.N) break;
          if (g.bias) x[j] += __ldg(g.bias + n + j);
          if (g.act == kActRelu) x[j] = fmaxf(x[j], 0.f);
          else if (g.act == kActClamp) x[j] = fminf(fmaxf(x[j], g.lo), g.hi);
          else if (g.act == kActSigmoid) x[j] = 1.0f / (1.0f + expf(-x[j]));
          if (g.mask) x[j] = g.mask[(size_t)m * g.ldmask + n + j] > 0.f ? x[j] : 0.f;
          if (g.accumulate) x[j] += dst[j];
        }
      }
      if (full4) *reinterpret_cast<float4*>(dst) = make_float4(x[0], x[1], x[2], x[3]);
      else
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < g.N) dst[j] = x[j];
    }
  } else {
    // ================================================================== MMA issue (warp-uniform, one elected lane)
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t idesc = umma_idesc_tf32(kGtM, kGtN);
    uint32_t full_par = 0;
    for (int c = 0; c < n_chunks; ++c) {
      const int s = c % kGtStages;
      mbar_wait(&S.full[s], (full_par >> s) & 1u);
      full_par ^= 1u << s;
      fence_after_sync();
      if (elect_one()) {
        const uint64_t dah = umma_smem_desc(smem_u32(S.a[s][0]), 128, kGtK * 32), dal = umma_smem_desc(smem_u32(S.a[s][1]), 128, kGtK * 32);
        const uint64_t dbh = umma_smem_desc(smem_u32(S.b[s][0]), 128, kGtK * 32), dbl = umma_smem_desc(smem_u32(S.b[s][1]), 128, kGtK * 32);
#pragma unroll
        for (int ks = 0; ks < kGtK / 8; ++ks) {
          const uint64_t adv = (uint64_t)((ks * 256) >> 4);
          mma_tf32_ss(tmem_u, dah + adv, dbh + adv, idesc, (c == 0 && ks == 0) ? 0u : 1u);
          mma_tf32_ss(tmem_u, dal + adv, dbh + adv, idesc, 1u);
          mma_tf32_ss(tmem_u, dah + adv, dbl + adv, idesc, 1u);
        }
        mma_commit(&S.empty[s]);
        if (c == n_chunks - 1) mma_commit(&S.done);
      }
      __syncwarp();
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == kGtLoadThreads / 32) tmem_dealloc(tmem, kGtN);
}

inline size_t tc_gemm_smem_bytes() {
  static_assert(sizeof(float) * kGtM * kGtTileLd <= sizeof(float) * 4 * kGtStages * kGtTileFloats, "the epilogue tile borrows the operand stages");
  return sizeof(GtSmem) + 1024;
}

// K splits for the tensor-core kernel (sgemm.cuh: gemm_splits_tc), bounded by the caller's workspace
inline int tc_gemm_splits(int M, int N, int K, size_t ws_floats) {
  int s = gemm_splits_tc(M, N, K);          // one full wave at two CTAs per SM (19 splits of a 2 x 2-tile gradient left half the chip idle)
  const size_t fit = ws_floats / ((size_t)M * N);
  if ((size_t)s > fit) s = (int)fit;
  return s < 1 ? 1 : s;
}

// Worth the tensor cores: enough work to fill 128 x 128 tiles.  (MACJD_TC_GEMM=0 keeps every GEMM on the FP32 kernel.)
inline bool tc_gemm_wanted(int M, int N, int K) {
  static const bool on = [] { const char* e = getenv("MACJD_TC_GEMM"); return !(e && e[0] == '0'); }();
  // (narrow outputs included: a [256 x 6] weight gradient over 200 k rows or a 5-column layer over 200 k rows is bound by
  // reading its tall operand, which the 128-row tiles with several chunks in flight do at 3-4x the FP32 kernel's rate)
  static const double min_macs = [] { const char* e = getenv("MACJD_TC_GEMM_MIN_LOG2"); return ldexp(1.0, e ? atoi(e) : 24); }();
  return on && M >= 64 && (double)M * N * K >= min_macs;
}

// returns false if the launch could not be set up (caller falls back to the FP32 kernel)
inline bool tc_gemm_launch(cudaStream_t st, int device, GemmArgs g, const GemmOpts& o) {
  static PerDeviceMax opted;
  const size_t smem = tc_gemm_smem_bytes();
  if (!opted.covers(device, smem)) {
    if (cudaFuncSetAttribute(tc_gemm_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(tc_gemm_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(tc_gemm_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(tc_gemm_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(tc_gemm_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(tc_gemm_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    opted.record(device, smem);
  }
  if (watchdog_arm(device) != MACJD_OK) return false;
  int splits = 1;
  if (o.splitk_ws && !o.bias && o.act == kActNone && !o.mask) splits = tc_gemm_splits(g.M, g.N, g.K, o.splitk_ws_floats);
  g.k_per_split = ((g.K + splits - 1) / splits + kGtK - 1) / kGtK * kGtK;
  if (g.k_per_split < kGtK) g.k_per_split = kGtK;
  splits = (g.K + g.k_per_split - 1) / g.k_per_split;
  if (splits < 1) splits = 1;
  g.partial = o.splitk_ws;
  const dim3 grid((g.N + kGtN - 1) / kGtN, (g.M + kGtM - 1) / kGtM, splits);
  // A(m,k) is k-contiguous unless ta; B(k,n) is k-contiguous (per output column n) when tb
  const bool ak = !g.ta, bk = g.tb != 0;
  // tall products: split B once (tc_pack_b_kernel), then every CTA fetches its B stages with bulk copies
  if (o.bpack_ws && g.M >= 8 * kGtM && tc_pack_b_floats(g.N, g.K) <= o.bpack_ws_floats &&
      (reinterpret_cast<uintptr_t>(o.bpack_ws) & 15) == 0) {
    const int kchunks = (g.K + kGtK - 1) / kGtK;
    MACJD_LAUNCH(tc_pack_b_kernel, dim3(grid.x * kchunks), dim3(256), 0, st, g.B, g.ldb, g.tb, g.N, g.K, o.bpack_ws);
    g.bpack = o.bpack_ws;
    g.bpack_kchunks = kchunks;
    if (ak) MACJD_LAUNCH((tc_gemm_kernel<true, true, true>), grid, dim3(kGtThreads), smem, st, g);
    else MACJD_LAUNCH((tc_gemm_kernel<false, true, true>), grid, dim3(kGtThreads), smem, st, g);
  } else
  if (ak && bk) MACJD_LAUNCH((tc_gemm_kernel<true, true>), grid, dim3(kGtThreads), smem, st, g);
  else if (ak) MACJD_LAUNCH((tc_gemm_kernel<true, false>), grid, dim3(kGtThreads), smem, st, g);
  else if (bk) MACJD_LAUNCH((tc_gemm_kernel<false, true>), grid, dim3(kGtThreads), smem, st, g);
  else MACJD_LAUNCH((tc_gemm_kernel<false, false>), grid, dim3(kGtThreads), smem, st, g);
  if (splits > 1) {
    const int total = g.M * g.N;
    MACJD_LAUNCH(splitk_reduce_kernel, dim3((total + 255) / 256), dim3(256), 0, st, (const float*)o.splitk_ws, splits, g.M, g.N,
                 g.C, g.ldc, o.accumulate);
  }
  return true;
}

}  // namespace tc
}  // namespace macjd
#endif  // !MACJD_TEST_HOST_EMULATION

namespace macjd {
// C = epi(A B): the tensor-core kernel where the problem fills its tiles, else the FP32 SIMT kernel
inline void gemm(cudaStream_t st, const float* A, int lda, bool ta, const float* B, int ldb, bool tb, float* C,
                 int ldc, int M, int N, int K, const GemmOpts& o = GemmOpts()) {
  if (M <= 0 || N <= 0) return;
#ifndef MACJD_TEST_HOST_EMULATION
  if (tc::tc_gemm_wanted(M, N, K)) {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && tc::tc_gemm_launch(st, dev, gemm_args(A, lda, ta, B, ldb, tb, C, ldc, M, N, K, o), o)) return;
  }
#endif
  gemm_simt(st, A, lda, ta, B, ldb, tb, C, ldc, M, N, K, o);
}
}  // namespace macjd
