// Time-unrolled agent forward (core/qmix.py:217-280: the learner's eval / target unrolls) as batched GEMMs, for network
// widths the fused CTA-pair kernel does not take (rnn_hidden_dim = 256 of BASELINE config 4: its operand tiles do not
// fit one SM's shared memory).  Only h_t -> h_t+1 is sequential; everything else is a dense layer over all T x M rows:
//   xf   = relu(obs fc1^T + b)                    [T M][H]        one GEMM
//   gx   = xf W_i{r,z,n}^T                        [T M][3H]       two GEMMs (r|z block, n block)
//   per step t:  gh = h_{t-1} W_h{r,z,n}^T  [M][3H] (two GEMMs), then the gate kernel -> h_t  (core/networks.py:88-114)
//   a1, a2 = actor hidden layers, P = sigmoid(a2 actor.4^T + b)   (core/networks.py:116-129)
//   qh   = h_t q.0[:, :H]^T + b for all rows, then Q(s, a, P_a) for every action, arg-max and gathers
//          (core/networks.py:131-180, core/qmix.py:141-147)
// The GEMMs go through gemm() (tc_gemm.cuh): tcgen05 3xTF32 when they fill its tiles, FP32 SIMT otherwise -- FP32-level
// accuracy either way.  Weights are the packed K-major copies of macjd_agent_weights ([in][out]).
#pragma once
#include "agent_act.cuh"
#include "tc_gemm.cuh"
#include "gru_rec_tc2.cuh"

namespace macjd {

// h' = GRU gates on pre-computed input / recurrent products (gate order r, z, n; core/networks.py:88-114 -> nn.GRUCell)
__global__ void __launch_bounds__(256) gru_gates_kernel(const float* __restrict__ gx, const float* __restrict__ gh,
                                                        const float* __restrict__ h_prev, const float* __restrict__ brz,
                                                        const float* __restrict__ bin, const float* __restrict__ bhn, int M, int H,
                                                        float* __restrict__ h_out) {
  grid_dependency_sync();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * H) return;
  const int row = idx / H, u = idx - row * H;
  const size_t g3 = (size_t)row * 3 * H;
  const float hr = gh ? gh[g3 + u] : 0.f, hz = gh ? gh[g3 + H + u] : 0.f, hn = gh ? gh[g3 + 2 * H + u] : 0.f;
  const float r = sigmoidf_ref(gx[g3 + u] + hr + brz[u]);
  const float z = sigmoidf_ref(gx[g3 + H + u] + hz + brz[H + u]);
  const float n = tanhf(gx[g3 + 2 * H + u] + bin[u] + r * (hn + bhn[u]));
  const float hp = h_prev ? h_prev[(size_t)row * H + u] : 0.f;
  h_out[(size_t)row * H + u] = (1.0f - z) * n + z * hp;
}

// Q(s, a, P_a) = w2 . relu(qh + W1[:, H + a] + P_a W1[:, H + A]) + b2 for every action of every row; arg-max
// (first maximum, no availability mask: the learner's double-DQN action) and the gather at sel_actions.
// One warp per row, lanes along the hidden units.
__global__ void __launch_bounds__(256) qhead_all_kernel(const float* __restrict__ qh, const float* __restrict__ P,
                                                        const float* __restrict__ w1a, const float* __restrict__ w1p,
                                                        const float* __restrict__ w2, const float* __restrict__ bq2, int R, int H, int A,
                                                        float* __restrict__ q_all, int* __restrict__ greedy,
                                                        const int* __restrict__ sel, float* __restrict__ q_sel) {
  grid_dependency_sync();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= R) return;
  const float b2 = bq2[0];
  float best = -INFINITY;
  int bi = 0;
  const int s_act = sel ? min(max(sel[row], 0), A - 1) : -1;
  float qs = 0.f;
  for (int a = 0; a < A; ++a) {
    const float pa = P[(size_t)row * A + a];
    float acc = 0.f;
    for (int u = lane; u < H; u += 32)
      acc = fmaf(w2[u], fmaxf(fmaf(pa, w1p[u], qh[(size_t)row * H + u] + w1a[(size_t)a * H + u]), 0.f), acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    const float q = acc + b2;
    if (lane == 0 && q_all) q_all[(size_t)row * A + a] = q;
    if (q > best) { best = q; bi = a; }
    if (a == s_act) qs = q;
  }
  if (lane == 0) {
    if (greedy) greedy[row] = bi;
    if (sel && q_sel) q_sel[row] = qs;
  }
}

// MACJD_REC_KERNEL=0 keeps the recurrence on the per-timestep GEMM + gate launches (comparison runs, tests)
inline bool rec_kernel_wanted() {
  const char* e = getenv("MACJD_REC_KERNEL");      // read per call (tens of ns): tests switch it
  return !(e && e[0] == '0');
}

struct UnrollWs {
  float *xf, *gx, *gh, *a1, *a2, *qh, *P, *hs, *bpack;
  size_t total, bpack_floats;
};
inline UnrollWs unroll_ws_layout(const macjd_agent_weights& w, int64_t M, int64_t T, float* base) {
  const int64_t rows = T * M, H = w.hidden, AH = w.actor_hidden, A = w.n_actions;
  UnrollWs u;
  size_t off = 0;
  auto take = [&](size_t n) { float* p = base ? base + off : nullptr; off += (n + 3) & ~(size_t)3; return p; };
  u.xf = take(rows * H); u.gx = take(rows * 3 * H); u.gh = take(M * 3 * H);
  u.a1 = take(rows * AH); u.a2 = take(rows * AH); u.qh = take(rows * H); u.P = take(rows * A); u.hs = take(rows * H);
  // the weights of the layer in flight, pre-split for the tensor-core GEMM (tc_gemm.cuh: tc_pack_b_kernel); the layers run
  // one after the other on the stream, so one region sized for the largest of them is enough
  u.bpack_floats = 0;
#ifndef MACJD_TEST_HOST_EMULATION
  {
    const int64_t O = w.obs_dim;
    const int64_t shapes[6][2] = {{3 * H, H}, {H, H}, {AH, AH}, {H, O}, {AH, O}, {A, AH}};     // (N, K) of every layer
    for (auto& s_ : shapes) {
      const size_t f = tc::tc_pack_b_floats((int)s_[0], (int)s_[1]);
      if (f > u.bpack_floats) u.bpack_floats = f;
    }
  }
#endif
  u.bpack = take(u.bpack_floats);
  u.total = off;
  return u;
}

inline int agent_unroll_gemm(const macjd_ctx* ctx, const macjd_agent_weights& W, const macjd_agent_io& io, float* ws_base,
                             size_t ws_floats) {
  const int M = io.n_rows, T = io.n_steps, O = W.obs_dim, H = W.hidden, AH = W.actor_hidden, A = W.n_actions;
  if (M < 0 || T < 1 || !io.obs) return MACJD_ERR_INVALID_ARG;
  if (io.actions || io.power || io.q_chosen || io.obs_group > 1 || io.part != 0) return MACJD_ERR_UNSUPPORTED;   // (no selection: acting is macjd_agent_forward)
  if (M == 0) return MACJD_OK;
  const UnrollWs u = unroll_ws_layout(W, M, T, ws_base);
  if (!ws_base || ws_floats < u.total) return MACJD_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)ctx->stream;
  const int rows = T * M;
  float* hs = io.hidden_seq ? io.hidden_seq : u.hs;
  GemmOpts relu;
  relu.act = kActRelu;
  relu.bpack_ws = u.bpack; relu.bpack_ws_floats = u.bpack_floats;
  // ---- input side of the GRU for all rows
  relu.bias = W.bfc1;
  gemm(st, io.obs, O, false, W.wfc1t, H, false, u.xf, H, rows, H, O, relu);
  bool use_rec = false;
#ifndef MACJD_TEST_HOST_EMULATION
  use_rec = W.rec_chunks && W.bgx && W.wiht && tc::gru_rec_supported(H) && rec_kernel_wanted();
#endif
  if (use_rec) {                      // the recurrence launch expects the input-side biases inside gate_x
    GemmOpts gb;
    gb.bias = W.bgx;
    gb.bpack_ws = u.bpack; gb.bpack_ws_floats = u.bpack_floats;
    gemm(st, u.xf, H, false, W.wiht, 3 * H, false, u.gx, 3 * H, rows, 3 * H, H, gb);
  } else if (W.wiht) {
    gemm(st, u.xf, H, false, W.wiht, 3 * H, false, u.gx, 3 * H, rows, 3 * H, H);
  } else {
    gemm(st, u.xf, H, false, W.wrzt, 2 * H, false, u.gx, 3 * H, rows, 2 * H, H);
    gemm(st, u.xf, H, false, W.wint, H, false, u.gx + 2 * H, 3 * H, rows, H, H);
  }
  // ---- the recurrence
  const float* h_first = (io.hidden_zero_init || !(io.hidden_in || io.hidden)) ? nullptr : (io.hidden_in ? io.hidden_in : io.hidden);
  const int gblocks = (M * H + 255) / 256;
  bool rec_done = false;
#ifndef MACJD_TEST_HOST_EMULATION
  if (use_rec) {
    // the whole recurrence as one launch: CTA pairs keep their rows' h in shared memory and stream W_hh (gru_rec_tc2.cuh)
    tc::RecArgs ra;
    ra.gate_x = u.gx; ra.h0 = h_first; ra.hidden_seq = hs; ra.hidden_out = nullptr;
    ra.bhn = W.bhn; ra.M = M; ra.T = T;
    const int rc = tc::gru_rec_launch(ctx, H, ra, W.rec_chunks);
    if (rc != MACJD_OK) return rc;
    rec_done = true;
  }
#endif
  for (int t = 0; t < T && !rec_done; ++t) {
    const float* h_prev = t == 0 ? h_first : hs + (size_t)(t - 1) * M * H;
    const float* gh = nullptr;
    if (h_prev) {
      if (W.whht) {
        gemm(st, h_prev, H, false, W.whht, 3 * H, false, u.gh, 3 * H, M, 3 * H, H);
      } else {
        gemm(st, h_prev, H, false, W.wrzt + (size_t)H * 2 * H, 2 * H, false, u.gh, 3 * H, M, 2 * H, H);
        gemm(st, h_prev, H, false, W.whnt, H, false, u.gh + 2 * H, 3 * H, M, H, H);
      }
      gh = u.gh;
    }
    MACJD_LAUNCH(gru_gates_kernel, dim3(gblocks), dim3(256), 0, st, (const float*)(u.gx + (size_t)t * M * 3 * H), gh, h_prev, W.brz,
                 W.bin, W.bhn, M, H, hs + (size_t)t * M * H);
  }
  if (io.hidden && cudaMemcpyAsync(io.hidden, hs + (size_t)(T - 1) * M * H, sizeof(float) * M * H, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
    return MACJD_ERR_CUDA;
  // ---- heads for all rows
  const bool want_heads = io.q_all || io.params_all || io.greedy || (io.sel_actions && io.q_sel);
  if (!want_heads) return MACJD_OK;
  float* P = io.params_all ? io.params_all : u.P;
  relu.bias = W.ba1;
  gemm(st, io.obs, O, false, W.wa1t, AH, false, u.a1, AH, rows, AH, O, relu);
  relu.bias = W.ba2;
  gemm(st, u.a1, AH, false, W.wa2t, AH, false, u.a2, AH, rows, AH, AH, relu);
  GemmOpts sig;
  sig.bpack_ws = u.bpack; sig.bpack_ws_floats = u.bpack_floats;
  sig.act = kActSigmoid;
  sig.bias = W.ba3;
  gemm(st, u.a2, AH, false, W.wa3t, A, false, P, A, rows, A, AH, sig);
  if (io.q_all || io.greedy || (io.sel_actions && io.q_sel)) {
    GemmOpts b;
    b.bias = W.bq1;
    b.bpack_ws = u.bpack; b.bpack_ws_floats = u.bpack_floats;
    gemm(st, hs, H, false, W.wqt, H, false, u.qh, H, rows, H, H, b);
    MACJD_LAUNCH(qhead_all_kernel, dim3((rows + 7) / 8), dim3(256), 0, st, (const float*)u.qh, (const float*)P, W.w1a, W.w1p, W.w2, W.bq2,
                 rows, H, A, io.q_all, (int*)io.greedy, (const int*)io.sel_actions, io.q_sel);
  }
  return MACJD_OK;
}

}  // namespace macjd
