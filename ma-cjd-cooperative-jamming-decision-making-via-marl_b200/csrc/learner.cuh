// Host-side composition of the learner kernels (enqueue only; no sync, no allocation).
//   mixer forward / backward    core/networks.py:250-316 and its autograd backward
//   Q-head forward / backward   core/networks.py:131-180 on stored hidden states (qmix.py:161-184)
//   TD target + masked loss     core/qmix.py:155,191-194
//   grad norm, clip, Adam       core/qmix.py:197-200
#pragma once
#include "learner_kernels.cuh"
#include "tc_gemm.cuh"

namespace macjd {

struct MixerWs {
  float *sn, *xhat, *h1, *w1, *hf, *wf, *b1, *hv, *v, *hidden;   // forward intermediates
  float *d_w1, *d_b1, *d_wf, *d_v, *d_h, *d_sn;                   // backward deltas
  float *splitk; size_t splitk_floats;
  float *colsum;
  float *bpack; size_t bpack_floats;    // the weights of the product in flight, pre-split for the tensor-core GEMM (tc_gemm.cuh)
  size_t total;
};

inline MixerWs mixer_ws_layout(const macjd_mixer_dims& d, float* base) {
  const size_t R = d.n_rows, S = d.state_dim, NE = (size_t)d.n_agents * d.embed_dim, E = d.embed_dim, HH = d.hyper_hidden;
  MixerWs w;
  size_t off = 0;
  auto take = [&](size_t n) { float* p = base ? base + off : nullptr; off += (n + 3) & ~(size_t)3; return p; };
  w.sn = take(R * S); w.xhat = take(R * S); w.h1 = take(R * HH); w.w1 = take(R * NE); w.hf = take(R * HH);
  w.wf = take(R * E); w.b1 = take(R * E); w.hv = take(R * E); w.v = take(R); w.hidden = take(R * E);
  w.d_w1 = take(R * NE); w.d_b1 = take(R * E); w.d_wf = take(R * E); w.d_v = take(R);
  w.d_h = take(R * (HH > E ? HH : E)); w.d_sn = take(R * S);
  size_t sk = 0;
  auto upd = [&](int M, int N) { const size_t f = gemm_splitk_ws_floats(M, N, (int)R); if (f > sk) sk = f; };
  upd((int)NE, (int)HH); upd((int)HH, (int)S); upd((int)E, (int)HH); upd((int)E, (int)S); upd(1, (int)E);
  w.splitk_floats = sk;
  w.splitk = take(sk);
  const size_t widest = NE > HH ? NE : HH;
  w.colsum = take(colsum_ws_floats((int)R, (int)(widest > S ? widest : S)));
  w.bpack_floats = 0;
#ifndef MACJD_TEST_HOST_EMULATION
  {
    // (N, K) of every product whose B operand is a weight matrix: the forward layers and the backward data products
    const size_t shapes[10][2] = {{HH, S}, {NE, HH}, {E, HH}, {E, S}, {1, E}, {HH, NE}, {S, HH}, {HH, E}, {S, E}, {E, 1}};
    for (auto& s_ : shapes) {
      const size_t f = tc::tc_pack_b_floats((int)s_[0], (int)s_[1]);
      if (f > w.bpack_floats) w.bpack_floats = f;
    }
  }
#endif
  w.bpack = take(w.bpack_floats);
  w.total = off;
  return w;
}

inline int rows_grid(int R) { return (R + 7) / 8; }   // 256 threads = 8 warps = 8 rows per block

inline int mixer_forward(cudaStream_t st, const macjd_mixer_dims& d, const macjd_mixer_params& p, const float* q,
                         const float* states, float* q_tot, float* ws_base, size_t ws_floats) {
  const int R = d.n_rows, S = d.state_dim, N = d.n_agents, E = d.embed_dim, HH = d.hyper_hidden, NE = N * E;
  if (R == 0) return MACJD_OK;
  const MixerWs w = mixer_ws_layout(d, ws_base);
  if (!ws_base || ws_floats < w.total) return MACJD_ERR_WORKSPACE;
  MACJD_LAUNCH(layernorm_fwd_kernel, dim3(rows_grid(R)), dim3(256), 0, st, states, R, S, (const float*)p.ln_w,
               (const float*)p.ln_b, w.sn, w.xhat);
  GemmOpts relu; relu.act = kActRelu;
  GemmOpts c05; c05.act = kActClamp; c05.lo = 0.f; c05.hi = 5.f;
  GemmOpts c55; c55.act = kActClamp; c55.lo = -5.f; c55.hi = 5.f;
  relu.bpack_ws = c05.bpack_ws = c55.bpack_ws = w.bpack;
  relu.bpack_ws_floats = c05.bpack_ws_floats = c55.bpack_ws_floats = w.bpack_floats;
  relu.bias = p.w1a_b; gemm(st, w.sn, S, false, p.w1a_w, S, true, w.h1, HH, R, HH, S, relu);
  c05.bias = p.w1b_b;  gemm(st, w.h1, HH, false, p.w1b_w, HH, true, w.w1, NE, R, NE, HH, c05);
  relu.bias = p.wfa_b; gemm(st, w.sn, S, false, p.wfa_w, S, true, w.hf, HH, R, HH, S, relu);
  c05.bias = p.wfb_b;  gemm(st, w.hf, HH, false, p.wfb_w, HH, true, w.wf, E, R, E, HH, c05);
  c55.bias = p.b1_b;   gemm(st, w.sn, S, false, p.b1_w, S, true, w.b1, E, R, E, S, c55);
  relu.bias = p.va_b;  gemm(st, w.sn, S, false, p.va_w, S, true, w.hv, E, R, E, S, relu);
  c55.bias = p.vb_b;   gemm(st, w.hv, E, false, p.vb_w, E, true, w.v, 1, R, 1, E, c55);
  MACJD_LAUNCH(mix_fwd_kernel, dim3(rows_grid(R)), dim3(256), 0, st, q, (const float*)w.w1, (const float*)w.b1,
               (const float*)w.wf, (const float*)w.v, R, N, E, w.hidden, q_tot);
  return MACJD_OK;
}

// Requires the workspace left by mixer_forward on the same rows.
inline int mixer_backward(cudaStream_t st, const macjd_mixer_dims& d, const macjd_mixer_params& p, const float* q,
                          const float* dq_tot, float* ws_base, size_t ws_floats, const macjd_mixer_params& g,
                          float* dq) {
  const int R = d.n_rows, S = d.state_dim, N = d.n_agents, E = d.embed_dim, HH = d.hyper_hidden, NE = N * E;
  if (R == 0) return MACJD_OK;
  const MixerWs w = mixer_ws_layout(d, ws_base);
  if (!ws_base || ws_floats < w.total) return MACJD_ERR_WORKSPACE;
  MACJD_LAUNCH(mix_bwd_kernel, dim3(rows_grid(R)), dim3(256), 0, st, dq_tot, q, (const float*)w.w1,
               (const float*)w.b1, (const float*)w.wf, (const float*)w.v, (const float*)w.hidden, R, N, E, w.d_w1,
               w.d_b1, w.d_wf, w.d_v, dq);
  GemmOpts sk; sk.splitk_ws = w.splitk; sk.splitk_ws_floats = w.splitk_floats;   // weight gradients: dY^T X
  GemmOpts acc; acc.accumulate = 1;
  GemmOpts plain;
  acc.bpack_ws = plain.bpack_ws = w.bpack;
  acc.bpack_ws_floats = plain.bpack_ws_floats = w.bpack_floats;
  // hyper_w_1 : sn -> relu(h1) -> w1
  gemm(st, w.d_w1, NE, true, w.h1, HH, false, g.w1b_w, HH, NE, HH, R, sk);
  colsum(st, w.d_w1, nullptr, R, NE, NE, g.w1b_b, w.colsum);
  { GemmOpts m = plain; m.mask = w.h1; m.ldmask = HH; gemm(st, w.d_w1, NE, false, p.w1b_w, HH, false, w.d_h, HH, R, HH, NE, m); }
  gemm(st, w.d_h, HH, true, w.sn, S, false, g.w1a_w, S, HH, S, R, sk);
  colsum(st, w.d_h, nullptr, R, HH, HH, g.w1a_b, w.colsum);
  gemm(st, w.d_h, HH, false, p.w1a_w, S, false, w.d_sn, S, R, S, HH, plain);
  // hyper_w_final : sn -> relu(hf) -> wf
  gemm(st, w.d_wf, E, true, w.hf, HH, false, g.wfb_w, HH, E, HH, R, sk);
  colsum(st, w.d_wf, nullptr, R, E, E, g.wfb_b, w.colsum);
  { GemmOpts m = plain; m.mask = w.hf; m.ldmask = HH; gemm(st, w.d_wf, E, false, p.wfb_w, HH, false, w.d_h, HH, R, HH, E, m); }
  gemm(st, w.d_h, HH, true, w.sn, S, false, g.wfa_w, S, HH, S, R, sk);
  colsum(st, w.d_h, nullptr, R, HH, HH, g.wfa_b, w.colsum);
  gemm(st, w.d_h, HH, false, p.wfa_w, S, false, w.d_sn, S, R, S, HH, acc);
  // hyper_b_1 : sn -> b1
  gemm(st, w.d_b1, E, true, w.sn, S, false, g.b1_w, S, E, S, R, sk);
  colsum(st, w.d_b1, nullptr, R, E, E, g.b1_b, w.colsum);
  gemm(st, w.d_b1, E, false, p.b1_w, S, false, w.d_sn, S, R, S, E, acc);
  // V : sn -> relu(hv) -> v
  gemm(st, w.d_v, 1, true, w.hv, E, false, g.vb_w, E, 1, E, R, sk);
  colsum(st, w.d_v, nullptr, R, 1, 1, g.vb_b, w.colsum);
  { GemmOpts m = plain; m.mask = w.hv; m.ldmask = E; gemm(st, w.d_v, 1, false, p.vb_w, E, false, w.d_h, E, R, E, 1, m); }
  gemm(st, w.d_h, E, true, w.sn, S, false, g.va_w, S, E, S, R, sk);
  colsum(st, w.d_h, nullptr, R, E, E, g.va_b, w.colsum);
  gemm(st, w.d_h, E, false, p.va_w, S, false, w.d_sn, S, R, S, E, acc);
  // LayerNorm affine parameters
  colsum(st, w.d_sn, w.xhat, R, S, S, g.ln_w, w.colsum);
  colsum(st, w.d_sn, nullptr, R, S, S, g.ln_b, w.colsum);
  return MACJD_OK;
}

inline size_t qhead_scratch_floats(const macjd_qhead_dims& d) {
  const size_t R = d.n_rows, A1 = d.n_actions + 1, H = d.hidden;
  size_t sk = gemm_splitk_ws_floats((int)H, (int)H, (int)R);
  const size_t sk2 = gemm_splitk_ws_floats((int)H, (int)A1, (int)R);
  if (sk2 > sk) sk = sk2;
  return ((R * A1 + 3) & ~(size_t)3) + ((sk + 3) & ~(size_t)3) + colsum_ws_floats((int)R, (int)H) + 16;
}

inline int qhead_forward(cudaStream_t st, const macjd_qhead_dims& d, const macjd_agent_weights& w, const float* hidden,
                         const int32_t* a_d, const float* a_c, float* q, float* hid, float* scratch, size_t scratch_floats) {
  const int R = d.n_rows, H = d.hidden, A = d.n_actions;
  if (R == 0) return MACJD_OK;
  GemmOpts o; o.bias = w.bq1;
  o.bpack_ws = scratch; o.bpack_ws_floats = scratch ? scratch_floats : 0;     // (optional: the pre-split weight matrix)
  gemm(st, hidden, H, false, w.wqt, H, false, hid, H, R, H, H, o);          // pre = h W1[:, :H]^T + b1
  MACJD_LAUNCH(qhead_tail_fwd_kernel, dim3(rows_grid(R)), dim3(256), 0, st, hid, (const int*)a_d, a_c, w.w1a, w.w1p,
               w.w2, w.bq2, R, H, A, q);
  return MACJD_OK;
}

inline int qhead_backward(cudaStream_t st, const macjd_qhead_dims& d, const macjd_agent_weights& w, const float* hidden,
                          const int32_t* a_d, const float* a_c, float* hid, const float* dq, float* g_w1, float* g_b1,
                          float* g_w2, float* g_b2, float* scratch, size_t scratch_floats) {
  const int R = d.n_rows, H = d.hidden, A = d.n_actions, ld = H + A + 1;
  if (R == 0) return MACJD_OK;
  if (!scratch || scratch_floats < qhead_scratch_floats(d)) return MACJD_ERR_WORKSPACE;
  float* xaug = scratch;
  float* skws = xaug + (((size_t)R * (A + 1) + 3) & ~(size_t)3);
  size_t skf = gemm_splitk_ws_floats(H, H, R);
  { const size_t f2 = gemm_splitk_ws_floats(H, A + 1, R); if (f2 > skf) skf = f2; }
  float* csws = skws + ((skf + 3) & ~(size_t)3);
  GemmOpts sk; sk.splitk_ws = skws; sk.splitk_ws_floats = skf;
  gemm(st, hid, H, true, dq, 1, false, g_w2, 1, H, 1, R, sk);               // dW2 = hid^T dq
  colsum(st, dq, nullptr, R, 1, 1, g_b2, csws);
  MACJD_LAUNCH(qhead_tail_bwd_kernel, dim3(rows_grid(R)), dim3(256), 0, st, hid, dq, (const int*)a_d, a_c, w.w2, R, H,
               A, xaug);
  gemm(st, hid, H, true, hidden, H, false, g_w1, ld, H, H, R, sk);          // dW1[:, :H]  = dhid^T h
  gemm(st, hid, H, true, xaug, A + 1, false, g_w1 + H, ld, H, A + 1, R, sk); // dW1[:, H:] = dhid^T [onehot, p]
  colsum(st, hid, nullptr, R, H, H, g_b1, csws);
  return MACJD_OK;
}

inline size_t td_scratch_floats(int R) { return (size_t)((R + kTdBlock - 1) / kTdBlock) * 4 + 4; }

inline int td_loss(cudaStream_t st, int R, const float* q_tot, const float* tq_tot, const float* reward,
                   const uint8_t* terminated, const uint8_t* filled, float gamma, float* dq_tot, float* targets,
                   float* sums, float* scratch, size_t scratch_floats) {
  if (R <= 0) return MACJD_ERR_INVALID_ARG;
  if (!scratch || scratch_floats < td_scratch_floats(R)) return MACJD_ERR_WORKSPACE;
  const int blocks = (R + kTdBlock - 1) / kTdBlock;
  MACJD_LAUNCH(td_partial_kernel, dim3(blocks), dim3(kTdBlock), 0, st, q_tot, tq_tot, reward, terminated, filled, gamma,
               R, dq_tot, targets, scratch);
  MACJD_LAUNCH(td_final_kernel, dim3(1), dim3(32), 0, st, (const float*)scratch, blocks, sums);
  return MACJD_OK;
}

constexpr int kSumsqBlocks = 64;
inline size_t opt_scratch_floats() { return kSumsqBlocks + 4; }

inline int clip_adam(cudaStream_t st, const macjd_opt_tensors& t, const float* grad, float* m, float* v,
                     const float* sums, float max_norm, float lr, float beta1, float beta2, float eps, int64_t step,
                     float* scal, float* scratch, size_t scratch_floats, const float* bias_corr = nullptr) {
  if (t.count < 1 || t.count > kMaxOptTensors || (step < 1 && !bias_corr)) return MACJD_ERR_INVALID_ARG;
  if (!scratch || scratch_floats < opt_scratch_floats()) return MACJD_ERR_WORKSPACE;
  OptTable tab;
  tab.count = t.count;
  int off = 0;
  for (int k = 0; k < t.count; ++k) {
    if (!t.param[k] || t.numel[k] < 0) return MACJD_ERR_INVALID_ARG;
    tab.param[k] = t.param[k];
    tab.offset[k] = off;
    off += (int)t.numel[k];
  }
  tab.offset[t.count] = off;
  MACJD_LAUNCH(sumsq_partial_kernel, dim3(kSumsqBlocks), dim3(256), 0, st, grad, off, scratch);
  MACJD_LAUNCH(clip_coef_kernel, dim3(1), dim3(32), 0, st, (const float*)scratch, kSumsqBlocks, sums, max_norm, scal);
  const double bc1 = bias_corr ? 1.0 : 1.0 - pow((double)beta1, (double)step);
  const double bc2 = bias_corr ? 1.0 : 1.0 - pow((double)beta2, (double)step);
  MACJD_LAUNCH(adam_kernel, dim3((off + 255) / 256), dim3(256), 0, st, tab, grad, m, v, (const float*)scal, lr, beta1,
               beta2, eps, (float)bc1, (float)sqrt(bc2), bias_corr);
  return MACJD_OK;
}

}  // namespace macjd
