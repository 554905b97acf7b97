// GRU recurrence of a time-unrolled agent pass for FEW rows (macjd_agent_forward, io.part == 4, H = 128).
//
// The learner unrolls both networks over whole episodes (core/qmix.py:129-147): M = B x n_agents rows -- 64 at the
// reference's batch of 32 episodes -- through T = 100 dependent steps.  On the CTA-pair tensor-core kernel that is ONE
// pair working through 100 x (144 MMAs + a gate epilogue): 10.2 us per step, 1.02 ms per unroll, half of a train step
// (profiles/r2_learner_c1.summary.csv), with 146 of 148 SMs idle.  Rows are independent, so this kernel splits them
// the other way: R rows per CTA (2 while that fits one wave of CTAs, else 4), and every CTA keeps ALL of rnn.weight_hh
// (3 x 128 x 128 FP32 = 192 KB) on chip for the whole launch.  Thread (u, kh) owns hidden unit u and half kh of the k
// range: of its 3 x 64 weights, 3 x kRecKReg live in REGISTERS (loaded once) and the rest in shared memory (k fastest,
// row stride padded so that a quarter-warp's 16-byte loads hit 32 distinct banks); the rows' states are read as
// broadcasts; 3 R accumulators, plain FP32 FMAs in k order; the two halves' partial sums meet through shared memory and
// each half applies the gates of half of the rows.  No tensor cores (an R x 384 x 128 product per step), no cluster, two
// __syncthreads per step (the state is double-buffered).  First version (128 threads, 4 rows, all weights in shared
// memory): 2.5 us per step on the B200 -- one warp per scheduler could not overlap its shared-memory reads (192 KB per
// step at 128 B / clk) with its FMAs.  Same gate arithmetic as the pair kernel's E4 (agent_act_tc2.cuh), on exact FP32
// products instead of 3xTF32 ones.
#pragma once
#include "macjd_common.cuh"
#include <stdlib.h>

namespace macjd {

constexpr int kRecH = 128;                    // hidden units
constexpr int kRecThreads = 2 * kRecH;        // (unit, k half)
constexpr int kRecKHalf = kRecH / 2;          // k per thread and gate
constexpr int kRecKReg = 32;                  // ... of which in registers
constexpr int kRecKSm = kRecKHalf - kRecKReg; // ... and in shared memory
constexpr int kRecWStride = 2 * kRecKSm + 4;  // floats per (gate, unit) row in shared memory (272 B: lanes u, u + 1 are 4 banks apart)
constexpr size_t rec_rows_smem_bytes(int R) { return sizeof(float) * (size_t)(3 * kRecH * kRecWStride + 2 * R * kRecH + 3 * R * kRecH); }

struct RecRowsArgs {
  const float* wrzt;        // [2H][2H]: rows H .. 2H-1 = rnn.weight_hh[0:2H]^T (columns: r gate, then z gate)
  const float* whnt;        // [H][H]   rnn.weight_hh[2H:3H]^T
  const float* brz;         // [2H] (bias_ih + bias_hh)[0:2H]
  const float* bin;         // [H]
  const float* bhn;         // [H]
  const float* gate_x;      // [T][M][3][H]: the input products W_ir xf, W_iz xf, W_in xf (no biases)
  const float* hidden_src;  // [M][H] initial state, or NULL = zeros
  float* hidden;            // [M][H] final state out, may be NULL
  float* hidden_seq;        // [T][M][H] every step's state out, may be NULL
  int M, T;
};

__device__ __forceinline__ float rec_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float rec_tanh(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }
__device__ __forceinline__ float rec_dot4(const float4 w, const float4 h, float acc) {
  return fmaf(w.w, h.w, fmaf(w.z, h.z, fmaf(w.y, h.y, fmaf(w.x, h.x, acc))));
}

template <int R>
__global__ void __launch_bounds__(kRecThreads, 1) gru_rec_rows_kernel(const RecRowsArgs a) {
  MACJD_DYNAMIC_SMEM(float, smem);
  constexpr int H = kRecH, WS = kRecWStride, KR = kRecKReg, KS = kRecKSm, RO = R / 2;   // RO: rows whose gates a half applies
  static_assert(R == 2 || R == 4, "rows per CTA");
  float* const Ws = smem;                          // [3][H][WS]: gate, unit, (k half, k)
  float* const hs = Ws + 3 * H * WS;               // [2][R][H]
  float* const red = hs + 2 * R * H;               // [R][3][H]: partial sums handed to the half that owns the row
  const int u = threadIdx.x & (H - 1), kh = threadIdx.x >> 7;     // (kh is warp-uniform)
  const int kb = kh * kRecKHalf;
  const int row0 = blockIdx.x * R;
  const int valid = min(R, a.M - row0);
  grid_dependency_wait();      // (launched as a programmatic dependent: the input pre-pass may still be running)

  // weights: K-major global rows (coalesced over u); this thread's first KR k's of each gate stay in registers
  float wreg[3][KR];
#pragma unroll
  for (int j = 0; j < KR; ++j) {
    const int k = kb + j;
    const float* rz = a.wrzt + (size_t)(H + k) * (2 * H);
    wreg[0][j] = __ldg(rz + u);
    wreg[1][j] = __ldg(rz + H + u);
    wreg[2][j] = __ldg(a.whnt + (size_t)k * H + u);
  }
#pragma unroll 4
  for (int j = 0; j < KS; ++j) {
    const int k = kb + KR + j;
    const float* rz = a.wrzt + (size_t)(H + k) * (2 * H);
    Ws[(0 * H + u) * WS + kh * KS + j] = __ldg(rz + u);
    Ws[(1 * H + u) * WS + kh * KS + j] = __ldg(rz + H + u);
    Ws[(2 * H + u) * WS + kh * KS + j] = __ldg(a.whnt + (size_t)k * H + u);
  }
  const float b_r = __ldg(a.brz + u), b_z = __ldg(a.brz + H + u), b_in = __ldg(a.bin + u), b_hn = __ldg(a.bhn + u);
  float hold[RO];                                  // the states of the rows this half owns: rows kh RO .. + RO - 1
#pragma unroll
  for (int i = 0; i < RO; ++i) {
    const int r = kh * RO + i;
    hold[i] = (a.hidden_src && r < valid) ? __ldg(a.hidden_src + (size_t)(row0 + r) * H + u) : 0.f;
    hs[r * H + u] = hold[i];
  }
  __syncthreads();

  const float4* const ws4 = reinterpret_cast<const float4*>(Ws + u * WS + kh * KS);     // + g * H * WS / 4 per gate
  for (int t = 0; t < a.T; ++t) {
    const float* const hcur = hs + (t & 1) * (R * H);
    float* const hnext = hs + ((t + 1) & 1) * (R * H);
    // this step's input products of the owned rows: requested before the products that hide their latency
    float gx[RO][3];
#pragma unroll
    for (int i = 0; i < RO; ++i) {
      const int r = kh * RO + i;
      const float* g = a.gate_x + ((size_t)t * a.M + row0 + r) * (3 * H) + u;
#pragma unroll
      for (int q = 0; q < 3; ++q) gx[i][q] = r < valid ? __ldg(g + q * H) : 0.f;
    }
    float acc[R][3];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r][0] = acc[r][1] = acc[r][2] = 0.f;
#pragma unroll
    for (int j4 = 0; j4 < KR / 4; ++j4) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float4 hv = *reinterpret_cast<const float4*>(hcur + r * H + kb + 4 * j4);
#pragma unroll
        for (int g = 0; g < 3; ++g)
          acc[r][g] = rec_dot4(make_float4(wreg[g][4 * j4], wreg[g][4 * j4 + 1], wreg[g][4 * j4 + 2], wreg[g][4 * j4 + 3]), hv, acc[r][g]);
      }
    }
#pragma unroll 4
    for (int j4 = 0; j4 < KS / 4; ++j4) {
      const float4 w0 = ws4[j4], w1 = ws4[(H * WS) / 4 + j4], w2 = ws4[2 * (H * WS) / 4 + j4];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float4 hv = *reinterpret_cast<const float4*>(hcur + r * H + kb + KR + 4 * j4);
        acc[r][0] = rec_dot4(w0, hv, acc[r][0]);
        acc[r][1] = rec_dot4(w1, hv, acc[r][1]);
        acc[r][2] = rec_dot4(w2, hv, acc[r][2]);
      }
    }
    // the partial sums of the rows the OTHER half owns go to it through shared memory
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (r / RO != kh) {
#pragma unroll
        for (int g = 0; g < 3; ++g) red[(r * 3 + g) * H + u] = acc[r][g];
      }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (r / RO == kh) {
        const int i = r % RO;
        const float sr = acc[r][0] + red[(r * 3 + 0) * H + u];
        const float sz = acc[r][1] + red[(r * 3 + 1) * H + u];
        const float sn = acc[r][2] + red[(r * 3 + 2) * H + u];
        const float rg = rec_sigmoid(sr + gx[i][0] + b_r);
        const float zg = rec_sigmoid(sz + gx[i][1] + b_z);
        const float n = rec_tanh(gx[i][2] + b_in + rg * (sn + b_hn));
        const float o = (1.0f - zg) * n + zg * hold[i];
        hold[i] = o;
        hnext[r * H + u] = o;
        if (r < valid) {
          const size_t off = (size_t)(row0 + r) * H + u;
          if (a.hidden_seq) a.hidden_seq[(size_t)t * a.M * H + off] = o;
          if (a.hidden && t == a.T - 1) a.hidden[off] = o;
        }
      }
    __syncthreads();
  }
}

// Rows up to which the recurrence of a part-4 call runs on this kernel (MACJD_REC_ROWS_MAX overrides; 0 = never).
// Default: two waves of 4-row CTAs on a B200 -- beyond that the CTA-pair
// tensor-core kernel, whose step costs the same 10 us for up to 74 x 128 rows, takes over.
inline int rec_rows_max() {
  const char* e = getenv("MACJD_REC_ROWS_MAX");            // read per call (tens of ns): tests switch it
  return e ? atoi(e) : 2 * kNumSMs * 4;
}

inline bool rec_rows_supported(const macjd_agent_weights& w, const macjd_agent_io& io) {
  return io.part == 4 && io.path != 1 && w.hidden == kRecH && w.wrzt && w.whnt && w.brz && w.bin && w.bhn && io.gate_x &&
         io.n_rows >= 1 && io.n_steps >= 1 && io.n_rows <= rec_rows_max();
}

inline int rec_rows_launch(const macjd_ctx* ctx, const macjd_agent_weights& w, const macjd_agent_io& io) {
  // 2 rows per CTA while those CTAs are one wave (the step is the same length, on twice the SMs), else 4
  const bool two = (io.n_rows + 1) / 2 <= kNumSMs;
  const size_t smem = rec_rows_smem_bytes(two ? 2 : 4);
#ifndef MACJD_TEST_HOST_EMULATION
  static PerDeviceMax opted;
  if (!opted.covers(ctx->device, rec_rows_smem_bytes(4))) {
    if (cudaFuncSetAttribute(gru_rec_rows_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rec_rows_smem_bytes(2)) != cudaSuccess ||
        cudaFuncSetAttribute(gru_rec_rows_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rec_rows_smem_bytes(4)) != cudaSuccess)
      return MACJD_ERR_CUDA;
    opted.record(ctx->device, rec_rows_smem_bytes(4));
  }
#endif
  RecRowsArgs a;
  a.wrzt = w.wrzt; a.whnt = w.whnt; a.brz = w.brz; a.bin = w.bin; a.bhn = w.bhn;
  a.gate_x = io.gate_x;
  const float* src = io.hidden_in ? io.hidden_in : io.hidden;      // initial state: a separate read-only source, or in place
  a.hidden_src = io.hidden_zero_init ? nullptr : src;
  a.hidden = io.hidden;
  a.hidden_seq = io.hidden_seq;
  a.M = io.n_rows;
  a.T = io.n_steps;
  const cudaStream_t st = (cudaStream_t)ctx->stream;
  if (two) MACJD_LAUNCH(gru_rec_rows_kernel<2>, (io.n_rows + 1) / 2, kRecThreads, smem, st, a);
  else MACJD_LAUNCH(gru_rec_rows_kernel<4>, (io.n_rows + 3) / 4, kRecThreads, smem, st, a);
  return MACJD_OK;
}

}  // namespace macjd
