// Re-layout of fc2_q_head after the optimiser step (core/qmix.py:197-200 updates it in place; it is the only part of
// the agent the learner trains, core/qmix.py:178): PyTorch-layout parameters -> the packed fields of
// macjd_agent_weights the kernels read (wqt, bq1, w1a, w1p, w2, bq2) and, where the tensor-core copy exists, the last
// H / kc weight chunks (q.0[:, :H] in the UMMA K-major layout, TF32 hi part then lo part) plus the Q-head entries of
// the constant block (q_c and the one-hot-column table).  ONE launch; on the host this was ~20 tensor operations per
// train step (0.25 ms of a 1.2 ms step at the reference batch, and 15 micro-kernels at the end of its critical path).
#pragma once
#include "macjd_common.cuh"

namespace macjd {

struct QheadRepackArgs {
  int H, A;
  const float* w1;        // [H][H + A + 1] fc2_q_head.0.weight
  const float* b1;        // [H]            fc2_q_head.0.bias
  const float* w2;        // [H]            fc2_q_head.2.weight
  const float* b2;        // [1]            fc2_q_head.2.bias
  float* wqt;             // [H][H]  w1[:, :H]^T
  float* bq1;             // [H]
  float* w1a;             // [A][H]  w1[:, H + a]
  float* w1p;             // [H]     w1[:, H + A]
  float* w2p;             // [H]
  float* bq2;             // [1]
  float* tc_chunks;       // optional (H == 128): the H / kc chunks of q.0[:, :H], [chunk][hi | lo][128 * kc]
  int kc;
  float* tc_q_c;          // optional: [128][4] = (bq1, w1p, w2, -) per unit
  float* tc_w1a;          // optional: [128][tc_w1a_stride], [unit][action] = w1[unit, H + action]
  int tc_w1a_stride;
};

__global__ void __launch_bounds__(256) qhead_repack_kernel(const QheadRepackArgs a) {
  grid_dependency_sync();
  const int H = a.H, A = a.A, ld = H + A + 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * H) return;
  const int u = i / H, k = i - u * H;                 // (unit = output row of w1, k = input column): coalesced reads
  const float v = a.w1[(size_t)u * ld + k];
  a.wqt[(size_t)k * H + u] = v;
  if (a.tc_chunks) {
    const int kc = a.kc, chunk = k / kc, kk = k - chunk * kc;
    const int within = (((u >> 3) * (kc >> 2) + (kk >> 2)) * 8 + (u & 7)) * 4 + (kk & 3);
    const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    float* dst = a.tc_chunks + (size_t)chunk * 2 * (128 * kc) + within;
    dst[0] = hi;
    dst[128 * kc] = v - hi;
  }
  if (k < A) {
    const float wa = a.w1[(size_t)u * ld + H + k];    // action k's one-hot column
    a.w1a[(size_t)k * H + u] = wa;
    if (a.tc_w1a) a.tc_w1a[(size_t)u * a.tc_w1a_stride + k] = wa;
  }
  if (k == 0) {
    const float b = a.b1[u], wp = a.w1[(size_t)u * ld + H + A], w2 = a.w2[u];
    a.bq1[u] = b;
    a.w1p[u] = wp;
    a.w2p[u] = w2;
    if (a.tc_q_c) { a.tc_q_c[4 * u + 0] = b; a.tc_q_c[4 * u + 1] = wp; a.tc_q_c[4 * u + 2] = w2; }
    if (u == 0) a.bq2[0] = a.b2[0];
  }
}

inline int qhead_repack(cudaStream_t st, const QheadRepackArgs& a) {
  if (a.H < 1 || a.A < 1 || a.A > a.H || !a.w1 || !a.b1 || !a.w2 || !a.b2 || !a.wqt || !a.bq1 || !a.w1a || !a.w1p || !a.w2p || !a.bq2)
    return MACJD_ERR_INVALID_ARG;
  if (a.tc_chunks && (a.H != 128 || a.kc < 4 || (a.kc & 3) || a.H % a.kc != 0)) return MACJD_ERR_INVALID_ARG;
  if ((a.tc_q_c || a.tc_w1a) && a.H != 128) return MACJD_ERR_INVALID_ARG;
  if (a.tc_w1a && a.tc_w1a_stride < a.A) return MACJD_ERR_INVALID_ARG;
  MACJD_LAUNCH(qhead_repack_kernel, (a.H * a.H + 255) / 256, 256, 0, st, a);
  return MACJD_OK;
}

}  // namespace macjd
