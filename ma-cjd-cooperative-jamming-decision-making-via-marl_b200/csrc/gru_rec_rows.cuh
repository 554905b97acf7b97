// GRU recurrence of a time-unrolled agent pass for FEW rows (macjd_agent_forward, io.part == 4, H = 128).
//
// The learner unrolls both networks over whole episodes (core/qmix.py:129-147): M = B x n_agents rows -- 64 at the
// reference's batch of 32 episodes -- through T = 100 dependent steps.  On the CTA-pair tensor-core kernel that is ONE
// pair working through 100 x (144 MMAs + a gate epilogue): 10.2 us per step, 1.02 ms per unroll, half of a train step
// (profiles/r2_learner_c1.summary.csv), with 146 of 148 SMs idle.  Rows are independent, so this kernel splits them
// the other way: kRecRows rows per CTA, every CTA keeps ALL of rnn.weight_hh (3 x 128 x 128 FP32 = 192 KB) in shared
// memory for the whole launch -- thread u owns hidden unit u: per step it reads its three weight rows with 16-byte
// loads (k fastest, row stride padded so that a quarter-warp's loads hit 32 distinct banks), the rows' states as
// broadcasts, and keeps 3 x kRecRows accumulators; plain FP32 FMAs in k order.  No tensor cores (a 4 x 384 x 128 product
// per step), no cluster, one __syncthreads per step (the state is double-buffered).  The step is bound by reading the
// weights out of shared memory (192 KB at 128 B / clk = 1.5 k cycles).  Same gate arithmetic as the pair kernel's E4
// (agent_act_tc2.cuh), on exact FP32 products instead of 3xTF32 ones.
#pragma once
#include "macjd_common.cuh"
#include <stdlib.h>

namespace macjd {

constexpr int kRecH = 128;                    // hidden units = threads per CTA
constexpr int kRecRows = 4;                   // rows per CTA
constexpr int kRecWStride = kRecH + 4;        // floats per weight row in shared memory (528 B: lanes u, u + 1 are 4 banks apart)
constexpr size_t kRecSmemBytes = sizeof(float) * (3 * kRecH * kRecWStride + 2 * kRecRows * kRecH);

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

__global__ void __launch_bounds__(kRecH, 1) gru_rec_rows_kernel(const RecRowsArgs a) {
  MACJD_DYNAMIC_SMEM(float, smem);
  constexpr int H = kRecH, WS = kRecWStride, R = kRecRows;
  float* const Ws = smem;                          // [3][H][WS]: gate g, unit u, k
  float* const hs = smem + 3 * H * WS;             // [2][R][H]
  const int u = threadIdx.x;
  const int row0 = blockIdx.x * R;
  const int valid = min(R, a.M - row0);
  grid_dependency_wait();      // (launched as a programmatic dependent: the input pre-pass may still be running)

  // weights: K-major global rows (coalesced over u) -> [gate][unit][k]
#pragma unroll 4
  for (int k = 0; k < H; ++k) {
    const float* rz = a.wrzt + (size_t)(H + k) * (2 * H);
    Ws[(0 * H + u) * WS + k] = __ldg(rz + u);
    Ws[(1 * H + u) * WS + k] = __ldg(rz + H + u);
    Ws[(2 * H + u) * WS + k] = __ldg(a.whnt + (size_t)k * H + u);
  }
  const float b_r = __ldg(a.brz + u), b_z = __ldg(a.brz + H + u), b_in = __ldg(a.bin + u), b_hn = __ldg(a.bhn + u);
  float hold[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    hold[r] = (a.hidden_src && r < valid) ? __ldg(a.hidden_src + (size_t)(row0 + r) * H + u) : 0.f;
    hs[r * H + u] = hold[r];
  }
  __syncthreads();

  const float4* const wr4 = reinterpret_cast<const float4*>(Ws + (0 * H + u) * WS);
  const float4* const wz4 = reinterpret_cast<const float4*>(Ws + (1 * H + u) * WS);
  const float4* const wn4 = reinterpret_cast<const float4*>(Ws + (2 * H + u) * WS);
  for (int t = 0; t < a.T; ++t) {
    const float* const hcur = hs + (t & 1) * (R * H);
    float* const hnext = hs + ((t + 1) & 1) * (R * H);
    // this step's input products: requested before the matrix-vector products that hide their latency
    float gx[R][3];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float* g = a.gate_x + ((size_t)t * a.M + row0 + r) * (3 * H) + u;
#pragma unroll
      for (int q = 0; q < 3; ++q) gx[r][q] = r < valid ? __ldg(g + q * H) : 0.f;
    }
    float ar[R], az[R], an[R];
#pragma unroll
    for (int r = 0; r < R; ++r) ar[r] = az[r] = an[r] = 0.f;
#pragma unroll 4
    for (int k4 = 0; k4 < H / 4; ++k4) {
      const float4 wr = wr4[k4], wz = wz4[k4], wn = wn4[k4];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float4 hv = *reinterpret_cast<const float4*>(hcur + r * H + 4 * k4);
        ar[r] = fmaf(wr.w, hv.w, fmaf(wr.z, hv.z, fmaf(wr.y, hv.y, fmaf(wr.x, hv.x, ar[r]))));
        az[r] = fmaf(wz.w, hv.w, fmaf(wz.z, hv.z, fmaf(wz.y, hv.y, fmaf(wz.x, hv.x, az[r]))));
        an[r] = fmaf(wn.w, hv.w, fmaf(wn.z, hv.z, fmaf(wn.y, hv.y, fmaf(wn.x, hv.x, an[r]))));
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float rg = rec_sigmoid(ar[r] + gx[r][0] + b_r);
      const float zg = rec_sigmoid(az[r] + gx[r][1] + b_z);
      const float n = rec_tanh(gx[r][2] + b_in + rg * (an[r] + b_hn));
      const float o = (1.0f - zg) * n + zg * hold[r];
      hold[r] = o;
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
// Default: two waves of CTAs on a B200 (2 x 148 x kRecRows rows, ~2 us per step) -- beyond that the CTA-pair
// tensor-core kernel, whose step costs the same 10 us for up to 74 x 128 rows, takes over.
inline int rec_rows_max() {
  const char* e = getenv("MACJD_REC_ROWS_MAX");            // read per call (tens of ns): tests switch it
  return e ? atoi(e) : 2 * 148 * kRecRows;
}

inline bool rec_rows_supported(const macjd_agent_weights& w, const macjd_agent_io& io) {
  return io.part == 4 && io.path != 1 && w.hidden == kRecH && w.wrzt && w.whnt && w.brz && w.bin && w.bhn && io.gate_x &&
         io.n_rows >= 1 && io.n_steps >= 1 && io.n_rows <= rec_rows_max();
}

inline int rec_rows_launch(const macjd_ctx* ctx, const macjd_agent_weights& w, const macjd_agent_io& io) {
#ifndef MACJD_TEST_HOST_EMULATION
  static PerDeviceMax opted;
  if (!opted.covers(ctx->device, kRecSmemBytes)) {
    if (cudaFuncSetAttribute(gru_rec_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRecSmemBytes) != cudaSuccess)
      return MACJD_ERR_CUDA;
    opted.record(ctx->device, kRecSmemBytes);
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
  const int grid = (io.n_rows + kRecRows - 1) / kRecRows;
  MACJD_LAUNCH(gru_rec_rows_kernel, grid, kRecH, kRecSmemBytes, (cudaStream_t)ctx->stream, a);
  return MACJD_OK;
}

}  // namespace macjd
