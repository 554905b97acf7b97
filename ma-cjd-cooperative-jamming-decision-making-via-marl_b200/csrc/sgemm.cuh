// General FP32 SIMT GEMM with fused epilogues, used by the learner kernels for the
// hypernetwork / Q-head layers (forward: X W^T, input gradient: dY W, weight gradient:
// dY^T X with a deterministic split over the row dimension).
//
//   C[m][n] (ldc) = epi( sum_k A(m,k) B(k,n) )      64x64 tile, 256 threads, 4x4 per thread
//   A(m,k) = ta ? A[k*lda + m] : A[m*lda + k]
//   B(k,n) = tb ? B[n*ldb + k] : B[k*ldb + n]        (tb = 1 is PyTorch's [out][in] weight)
//   epi: + bias[n]; relu | clamp(lo, hi); * (mask[m][n] > 0); += C (accumulate)
//
// Arbitrary M, N, K (bounds-checked scalar loads, mapped so that the contiguous dimension of
// each operand runs along the lanes).  grid.z > 1 splits K; partial tiles go to a workspace
// and a second kernel reduces them in a fixed order (bit-reproducible gradients).
#pragma once
#include "macjd_common.cuh"

namespace macjd {

enum GemmAct : int { kActNone = 0, kActRelu = 1, kActClamp = 2, kActSigmoid = 3 };

struct GemmArgs {
  const float* A; const float* B; float* C;
  int M, N, K, lda, ldb, ldc;
  int ta, tb;
  const float* bias;      // [N] or null
  int act; float lo, hi;
  const float* mask; int ldmask;   // multiply by (mask > 0), or null
  int accumulate;         // C += result
  int k_per_split;        // K range per blockIdx.z (multiple of 16)
  float* partial;         // [splits][M][N] when gridDim.z > 1
  const float* bpack;     // tensor-core kernel only: B pre-split and pre-laid-out (tc_gemm.cuh: tc_pack_b_kernel), or null
  int bpack_kchunks;      // 16-k chunks per 128-column tile in bpack
};

constexpr int kGM = 64, kGN = 64, kGK = 16;

__global__ void __launch_bounds__(256) sgemm_kernel(const GemmArgs g) {
  grid_dependency_sync();   // no-op unless launched as a programmatic dependent (MACJD_LAUNCH)
  __shared__ float As[kGK][kGM + 4];
  __shared__ float Bs[kGK][kGN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * kGM, n0 = blockIdx.x * kGN;
  const int kbeg = blockIdx.z * g.k_per_split;
  const int kend = min(g.K, kbeg + g.k_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += kGK) {
    // ---- stage A tile (64 x 16) and B tile (16 x 64), 4 elements per thread each
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = tid + q * 256;
      int m, k;
      if (g.ta) { m = e & 63; k = e >> 6; } else { k = e & 15; m = e >> 4; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < g.M && gk < kend) v = g.ta ? g.A[(size_t)gk * g.lda + gm] : g.A[(size_t)gm * g.lda + gk];
      As[k][m] = v;
      int n, kb;
      if (g.tb) { kb = e & 15; n = e >> 4; } else { n = e & 63; kb = e >> 6; }
      const int gn = n0 + n, gkb = k0 + kb;
      float w = 0.f;
      if (gn < g.N && gkb < kend) w = g.tb ? g.B[(size_t)gn * g.ldb + gkb] : g.B[(size_t)gkb * g.ldb + gn];
      Bs[kb][n] = w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kGK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (gridDim.z > 1) {
        g.partial[((size_t)blockIdx.z * g.M + m) * g.N + n] = v;
        continue;
      }
      if (g.bias) v += g.bias[n];
      if (g.act == kActRelu) v = fmaxf(v, 0.f);
      else if (g.act == kActClamp) v = fminf(fmaxf(v, g.lo), g.hi);
      else if (g.act == kActSigmoid) v = 1.0f / (1.0f + expf(-v));
      if (g.mask) v = g.mask[(size_t)m * g.ldmask + n] > 0.f ? v : 0.f;
      float* c = g.C + (size_t)m * g.ldc + n;
      *c = g.accumulate ? *c + v : v;
    }
  }
}

// C[m][n] = (accumulate ? C : 0) + sum_s partial[s][m][n]   (fixed order)
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* partial, int splits, int M, int N,
                                                            float* C, int ldc, int accumulate) {
  grid_dependency_sync();   // no-op unless launched as a programmatic dependent (MACJD_LAUNCH)
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * N) return;
  const int m = idx / N, n = idx - m * N;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += partial[(size_t)z * M * N + idx];
  float* c = C + (size_t)m * ldc + n;
  *c = accumulate ? *c + s : s;
}

struct GemmOpts {
  const float* bias = nullptr;
  int act = kActNone; float lo = 0.f, hi = 0.f;
  const float* mask = nullptr; int ldmask = 0;
  int accumulate = 0;
  float* splitk_ws = nullptr; size_t splitk_ws_floats = 0;   // enables split-K when useful
  float* bpack_ws = nullptr; size_t bpack_ws_floats = 0;     // tensor-core kernel: room to pre-split B once per call (tall products)
};

// Number of K splits used for a reduction over `K` rows into an M x N result.
inline int gemm_splits(int M, int N, int K) {
  const int tiles = ((M + kGM - 1) / kGM) * ((N + kGN - 1) / kGN);
  if (K < 2048 || tiles >= kNumSMs) return 1;
  int s = (2 * kNumSMs + tiles - 1) / tiles;
  // slices of >= 128 rows: the learner's weight gradients reduce a few thousand rows into one to four
  // tiles -- with 512-row slices that was 7 to 28 CTAs walking 28 k-steps each (54 us per GEMM)
  const int maxs = (K + 127) / 128;
  if (s > maxs) s = maxs;
  return s < 1 ? 1 : s;
}

// K splits of the tensor-core kernel (tc_gemm.cuh: 128 x 128 tiles, two CTAs per SM): enough CTAs for one full wave,
// slices of at least 256 rows
inline int gemm_splits_tc(int M, int N, int K) {
  const int tiles = ((M + 127) / 128) * ((N + 127) / 128);
  if (K < 2048 || tiles >= kNumSMs) return 1;
  int s = (2 * kNumSMs + tiles - 1) / tiles;
  const int maxs = (K + 255) / 256;
  if (s > maxs) s = maxs;
  return s < 1 ? 1 : s;
}

// workspace for either kernel's partial sums
inline size_t gemm_splitk_ws_floats(int M, int N, int K) {
  const int a = gemm_splits(M, N, K), b = gemm_splits_tc(M, N, K);
  const int s = a > b ? a : b;
  return s > 1 ? (size_t)s * M * N : 0;
}

inline GemmArgs gemm_args(const float* A, int lda, bool ta, const float* B, int ldb, bool tb, float* C, int ldc, int M, int N,
                          int K, const GemmOpts& o) {
  GemmArgs g;
  g.A = A; g.B = B; g.C = C; g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldb = ldb; g.ldc = ldc;
  g.ta = ta; g.tb = tb; g.bias = o.bias; g.act = o.act; g.lo = o.lo; g.hi = o.hi;
  g.mask = o.mask; g.ldmask = o.ldmask; g.accumulate = o.accumulate;
  g.k_per_split = K; g.partial = nullptr;
  g.bpack = nullptr; g.bpack_kchunks = 0;
  return g;
}

// the FP32 SIMT kernel (small problems, host emulation)
inline void gemm_simt(cudaStream_t st, const float* A, int lda, bool ta, const float* B, int ldb, bool tb, float* C,
                      int ldc, int M, int N, int K, const GemmOpts& o = GemmOpts()) {
  if (M <= 0 || N <= 0) return;
  GemmArgs g = gemm_args(A, lda, ta, B, ldb, tb, C, ldc, M, N, K, o);
  int splits = 1;
  if (o.splitk_ws && !o.bias && o.act == kActNone && !o.mask) {
    splits = gemm_splits(M, N, K);
    if ((size_t)splits * M * N > o.splitk_ws_floats) splits = 1;
  }
  g.k_per_split = ((K + splits - 1) / splits + kGK - 1) / kGK * kGK;
  if (g.k_per_split < kGK) g.k_per_split = kGK;
  splits = (K + g.k_per_split - 1) / g.k_per_split;
  if (splits < 1) splits = 1;
  g.partial = o.splitk_ws;
  dim3 grid((N + kGN - 1) / kGN, (M + kGM - 1) / kGM, splits);
  MACJD_LAUNCH(sgemm_kernel, grid, dim3(256), 0, st, g);
  if (splits > 1) {
    const int total = M * N;
    MACJD_LAUNCH(splitk_reduce_kernel, dim3((total + 255) / 256), dim3(256), 0, st,
                 (const float*)o.splitk_ws, splits, M, N, C, ldc, o.accumulate);
  }
}

}  // namespace macjd
