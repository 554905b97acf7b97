// Fused per-timestep agent step (K2-K6 of SURVEY.md), FP32 SIMT.
//
// One CTA owns a tile of TM agent rows (row = (env, agent)) and runs the whole network
// for them, optionally for T consecutive timesteps with the recurrent state kept in
// shared memory (the learner's eval / target unrolls, core/qmix.py:217-280):
//
//   a1 = relu(W_a1 x + b)   a2 = relu(W_a2 a1 + b)   P = sigmoid(W_a3 a2 + b)   actor,  networks.py:116-129
//   xf = relu(W_fc1 x + b)                                                      fc1,    networks.py:102
//   r,z = sigmoid(W_i{r,z} xf + W_h{r,z} h + b)   n = tanh(W_in xf + b_in + r (W_hn h + b_hn))
//   h'  = (1 - z) n + z h                                                       GRUCell, networks.py:113
//   pre = W1[:, :H] h' + b1        (shared by all actions -- the reference recomputes it A times)
//   Q_a = w2 . relu(pre + W1[:, H+a] + P_a W1[:, H+A]) + b2                     Q-head, networks.py:131-180
//   mask, first-max argmax, epsilon-greedy, gather P / Q                        mac.py:138-164, action_selectors.py
//
// Dense layers are register-tiled SIMT GEMMs: activations live K-major in shared memory
// ([k][row], row stride TM+4), weights are streamed from the packed K-major copy in L2 in
// 32x64 chunks through a 2-stage cp.async pipeline; 256 threads, each RT x 4 outputs.
// Bound: FP32 FMA pipe (AI ~ 350 FLOP/B).
#pragma once
#include "macjd_common.cuh"

namespace macjd {

constexpr int kAgentThreads = 256;
constexpr int kKC = 32;  // K rows per staged weight chunk
constexpr int kNC = 64;  // output columns per GEMM pass

struct AgentArgs {
  macjd_agent_weights w;
  macjd_agent_io io;
};

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
#ifdef MACJD_TEST_HOST_EMULATION
  *reinterpret_cast<float4*>(smem_dst) = *reinterpret_cast<const float4*>(gsrc);
#else
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
#endif
}
__device__ __forceinline__ void cp_async_commit() {
#ifndef MACJD_TEST_HOST_EMULATION
  asm volatile("cp.async.commit_group;\n" ::);
#endif
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
#ifndef MACJD_TEST_HOST_EMULATION
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
#endif
}

template <int RT>
__device__ __forceinline__ void zero_acc(float (&acc)[RT][4]) {
#pragma unroll
  for (int i = 0; i < RT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
}

// acc[RT][4] += As[K][TM]^T-tile x Wt[K][col0 + 64]: rows ty*RT.., cols col0 + tx*4..
// As: shared, K-major, row stride 16*RT+4.  Wt: global, K-major, leading dim ldw.
// K must be a multiple of 32.  Ends with a __syncthreads (stage + As are free again).
template <int RT>
__device__ __forceinline__ void gemm_pass(float (&acc)[RT][4], const float* As, int K,
                                          const float* __restrict__ Wt, int ldw, int col0, float* wst, int tid) {
  constexpr int TMp = 16 * RT + 4;
  const int tx = tid & 15, ty = tid >> 4;
  const int nchunks = K / kKC;
  const float* src0 = Wt + (size_t)ty * ldw + col0 + tx * 4;
  float* dst0 = wst + ty * kNC + tx * 4;
  // stage chunk 0
  cp_async16(dst0, src0);
  cp_async16(dst0 + 16 * kNC, src0 + (size_t)16 * ldw);
  cp_async_commit();
  for (int c = 0; c < nchunks; ++c) {
    if (c + 1 < nchunks) {
      const float* s = src0 + (size_t)(c + 1) * kKC * ldw;
      float* d = dst0 + ((c + 1) & 1) * (kKC * kNC);
      cp_async16(d, s);
      cp_async16(d + 16 * kNC, s + (size_t)16 * ldw);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* wb = wst + (c & 1) * (kKC * kNC) + tx * 4;
    const float* ab = As + (size_t)(c * kKC) * TMp + ty * RT;
#pragma unroll
    for (int kk = 0; kk < kKC; ++kk) {
      const float4 w = *reinterpret_cast<const float4*>(wb + kk * kNC);
      float a[RT];
      if (RT == 4) {
        const float4 av = *reinterpret_cast<const float4*>(ab + kk * TMp);
        a[0] = av.x; a[1] = av.y; a[2] = av.z; a[3] = av.w;
      } else {
        const float2 av = *reinterpret_cast<const float2*>(ab + kk * TMp);
        a[0] = av.x; a[1] = av.y;
      }
#pragma unroll
      for (int i = 0; i < RT; ++i) {
        acc[i][0] = fmaf(a[i], w.x, acc[i][0]);
        acc[i][1] = fmaf(a[i], w.y, acc[i][1]);
        acc[i][2] = fmaf(a[i], w.z, acc[i][2]);
        acc[i][3] = fmaf(a[i], w.w, acc[i][3]);
      }
    }
    __syncthreads();
  }
}

// out[col][row] = f(acc + bias[col]) into a K-major shared activation buffer.
template <int RT, bool RELU>
__device__ __forceinline__ void store_tile(const float (&acc)[RT][4], const float* __restrict__ bias, int col0,
                                           float* Os, int tid) {
  constexpr int TMp = 16 * RT + 4;
  const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int col = col0 + tx * 4 + j;
    const float b = __ldg(bias + col);
#pragma unroll
    for (int i = 0; i < RT; ++i) {
      float v = acc[i][j] + b;
      if (RELU) v = fmaxf(v, 0.f);
      Os[(size_t)col * TMp + ty * RT + i] = v;
    }
  }
}

// Dense layer over all output chunks: Os[N][TM] = f(As[K][TM]^T W + b)
template <int RT, bool RELU>
__device__ __forceinline__ void dense_layer(const float* As, int K, const float* __restrict__ Wt,
                                            const float* __restrict__ bias, int N, float* Os, float* wst, int tid) {
  for (int col0 = 0; col0 < N; col0 += kNC) {
    float acc[RT][4];
    zero_acc<RT>(acc);
    gemm_pass<RT>(acc, As, K, Wt, N, col0, wst, tid);
    store_tile<RT, RELU>(acc, bias, col0, Os, tid);
  }
}

template <int RT>
__global__ void __launch_bounds__(kAgentThreads, 1) agent_forward_kernel(const AgentArgs a) {
  constexpr int TM = 16 * RT, TMp = TM + 4;
  const macjd_agent_weights& W = a.w;
  const macjd_agent_io& io = a.io;
  const int O = W.obs_dim, Op = W.obs_pad, H = W.hidden, AH = W.actor_hidden, A = W.n_actions;
  const int M = io.n_rows, T = io.n_steps;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int row0 = blockIdx.x * TM;
  const int valid = min(TM, M - row0);
  const int HB = H > AH ? H : AH;

  MACJD_DYNAMIC_SMEM(float, smem);
  float* Xs = smem;                          // [Op][TMp]  observation tile
  float* B0 = Xs + (size_t)Op * TMp;         // [HB][TMp]  a1 -> xf -> pre
  float* B1 = B0 + (size_t)HB * TMp;         // [HB][TMp]  a2 -> h'
  float* B2 = B1 + (size_t)HB * TMp;         // [HB][TMp]  h (previous step); swaps role with B1 every step
  float* Ps = B2 + (size_t)HB * TMp;          // [A][TMp]   actor parameters
  float* Qs = Ps + (size_t)A * TMp;          // [A][TMp]   Q per action
  float* wst = Qs + (size_t)A * TMp;         // [2][32][64] weight stage

  // recurrent state tile: K-major [unit][row]; lanes run along `unit` (coalesced global)
  for (int idx = tid; idx < TM * H; idx += kAgentThreads) {
    const int r = idx / H, k = idx - r * H;
    float v = 0.f;
    if (r < valid && io.hidden && !io.hidden_zero_init) v = io.hidden[(size_t)(row0 + r) * H + k];
    B2[(size_t)k * TMp + r] = v;
  }

  for (int t = 0; t < T; ++t) {
    const size_t tM = (size_t)t * M;
    // ---- observation tile, zero padded to [Op][TM]
    {
      const float* obs = io.obs + (tM + row0) * O;
      for (int idx = tid; idx < TM * Op; idx += kAgentThreads) {
        const int r = idx / Op, k = idx - r * Op;
        float v = 0.f;
        if (r < valid && k < O) v = __ldg(obs + (size_t)r * O + k);
        Xs[(size_t)k * TMp + r] = v;
      }
    }
    // (the first __syncthreads inside gemm_pass orders these stores before any read)

    // ---- actor MLP
    dense_layer<RT, true>(Xs, Op, W.wa1t, W.ba1, AH, B0, wst, tid);
    dense_layer<RT, true>(B0, AH, W.wa2t, W.ba2, AH, B1, wst, tid);
    __syncthreads();
    for (int item = tid; item < A * TM; item += kAgentThreads) {
      const int act = item / TM, r = item - act * TM;
      float s = __ldg(W.ba3 + act);
      for (int k = 0; k < AH; ++k) s = fmaf(B1[(size_t)k * TMp + r], __ldg(W.wa3t + (size_t)k * A + act), s);
      Ps[act * TMp + r] = 1.0f / (1.0f + expf(-s));
    }
    // ---- fc1
    dense_layer<RT, true>(Xs, Op, W.wfc1t, W.bfc1, H, B0, wst, tid);
    __syncthreads();  // Ps complete (B1 may be overwritten), xf complete

    // ---- GRU cell, 64 hidden units at a time; gates combined in registers
    for (int u0 = 0; u0 < H; u0 += kNC) {
      float g[RT][4], rg[RT][4], zg[RT][4], hn[RT][4];
      zero_acc<RT>(g);
      gemm_pass<RT>(g, B0, H, W.wrzt, 2 * H, u0, wst, tid);
      gemm_pass<RT>(g, B2, H, W.wrzt + (size_t)H * 2 * H, 2 * H, u0, wst, tid);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float b = __ldg(W.brz + u0 + tx * 4 + j);
#pragma unroll
        for (int i = 0; i < RT; ++i) rg[i][j] = 1.0f / (1.0f + expf(-(g[i][j] + b)));
      }
      zero_acc<RT>(g);
      gemm_pass<RT>(g, B0, H, W.wrzt, 2 * H, H + u0, wst, tid);
      gemm_pass<RT>(g, B2, H, W.wrzt + (size_t)H * 2 * H, 2 * H, H + u0, wst, tid);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float b = __ldg(W.brz + H + u0 + tx * 4 + j);
#pragma unroll
        for (int i = 0; i < RT; ++i) zg[i][j] = 1.0f / (1.0f + expf(-(g[i][j] + b)));
      }
      zero_acc<RT>(g);
      gemm_pass<RT>(g, B0, H, W.wint, H, u0, wst, tid);
      zero_acc<RT>(hn);
      gemm_pass<RT>(hn, B2, H, W.whnt, H, u0, wst, tid);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int unit = u0 + tx * 4 + j;
        const float bi = __ldg(W.bin + unit), bh = __ldg(W.bhn + unit);
#pragma unroll
        for (int i = 0; i < RT; ++i) {
          const float n = tanhf(g[i][j] + bi + rg[i][j] * (hn[i][j] + bh));
          const float hold = B2[(size_t)unit * TMp + ty * RT + i];
          g[i][j] = (1.0f - zg[i][j]) * n + zg[i][j] * hold;
          B1[(size_t)unit * TMp + ty * RT + i] = g[i][j];
        }
      }
      // h' to global: 4 consecutive units per row -> 16-byte stores, 256 B per row per warp half
#pragma unroll
      for (int i = 0; i < RT; ++i) {
        const int r = ty * RT + i;
        if (r < valid) {
          const float4 v = make_float4(g[i][0], g[i][1], g[i][2], g[i][3]);
          const size_t off = (size_t)(row0 + r) * H + u0 + tx * 4;
          if (io.hidden_seq) *reinterpret_cast<float4*>(io.hidden_seq + tM * H + off) = v;
          if (io.hidden && t == T - 1) *reinterpret_cast<float4*>(io.hidden + off) = v;
        }
      }
    }

    // ---- Q-head: shared pre-activation, then the per-action tail
    dense_layer<RT, false>(B1, H, W.wqt, W.bq1, H, B0, wst, tid);
    __syncthreads();
    {
      const float bq2 = __ldg(W.bq2);
      for (int item = tid; item < A * TM; item += kAgentThreads) {
        const int act = item / TM, r = item - act * TM;
        const float p = Ps[act * TMp + r];
        const float* w1a = W.w1a + (size_t)act * H;
        float s = bq2;
        for (int n = 0; n < H; ++n) {
          const float v = B0[(size_t)n * TMp + r] + __ldg(w1a + n) + p * __ldg(W.w1p + n);
          s = fmaf(__ldg(W.w2 + n), fmaxf(v, 0.f), s);
        }
        Qs[act * TMp + r] = s;
      }
    }
    __syncthreads();

    // ---- outputs: dense [row][action] tables (coalesced), then one thread per row selects
    if (io.q_all || io.params_all) {
      for (int idx = tid; idx < valid * A; idx += kAgentThreads) {
        const int r = idx / A, act = idx - r * A;
        const size_t o = (tM + row0) * A + idx;
        if (io.q_all) io.q_all[o] = Qs[act * TMp + r];
        if (io.params_all) io.params_all[o] = Ps[act * TMp + r];
      }
    }
    if (tid < valid) {
      const int r = tid;
      const size_t m = tM + row0 + r;
      const uint8_t* av = io.avail ? io.avail + m * A : nullptr;
      float best = -INFINITY, bestm = -INFINITY;
      int bi = 0, bim = 0, n_avail = 0;
      for (int act = 0; act < A; ++act) {
        const float q = Qs[act * TMp + r];
        if (q > best) { best = q; bi = act; }                  // first max, no mask (qmix.py:143)
        const bool ok = av ? (av[act] != 0) : true;
        n_avail += ok ? 1 : 0;
        const float qm = ok ? q : -INFINITY;                   // mac.py:142
        if (qm > bestm) { bestm = qm; bim = act; }
      }
      if (io.greedy) io.greedy[m] = bi;
      if (io.sel_actions && io.q_sel) {
        int s = io.sel_actions[m];
        s = s < 0 ? 0 : (s >= A ? A - 1 : s);
        io.q_sel[m] = Qs[s * TMp + r];
      }
      if (io.actions) {
        int chosen = bim;
        if (!io.test_mode) {                                   // action_selectors.py:39-57
          const uint32_t row_id = (uint32_t)(row0 + r);
          const float u = io.u_eps ? io.u_eps[m] : philox_uniform(io.seed, kStreamEpsilon, row_id, io.rng_step + t, 0);
          if (u < io.epsilon) {
            if (io.rand_actions) {
              chosen = io.rand_actions[m];
            } else {
              // uniform over the available actions (multinomial over the 0/1 mask)
              const float u2 = philox_uniform(io.seed, kStreamRandomAction, row_id, io.rng_step + t, 0);
              const int navl = n_avail > 0 ? n_avail : A;
              int kth = (int)(u2 * (float)navl);
              kth = kth >= navl ? navl - 1 : kth;
              chosen = 0;
              for (int act = 0, seen = 0; act < A; ++act) {
                const bool ok = (n_avail == 0) || !av || av[act] != 0;
                if (ok) { if (seen == kth) { chosen = act; break; } ++seen; }
              }
            }
            chosen = chosen < 0 ? 0 : (chosen >= A ? A - 1 : chosen);
          }
        }
        io.actions[m] = chosen;
        if (io.power) io.power[m] = Ps[chosen * TMp + r];
        if (io.q_chosen) io.q_chosen[m] = Qs[chosen * TMp + r];
      }
    }
    __syncthreads();
    // h' becomes h for the next unrolled step
    float* tmp = B1; B1 = B2; B2 = tmp;
  }
}

inline size_t agent_smem_bytes(const macjd_agent_weights& w, int TM) {
  const int TMp = TM + 4;
  const int HB = w.hidden > w.actor_hidden ? w.hidden : w.actor_hidden;
  return sizeof(float) * ((size_t)w.obs_pad * TMp + (size_t)3 * HB * TMp +
                          (size_t)2 * w.n_actions * TMp + 2 * kKC * kNC);
}

inline int agent_launch(const macjd_ctx* ctx, const macjd_agent_weights* w, const macjd_agent_io* io) {
  if (!ctx || !w || !io) return MACJD_ERR_INVALID_ARG;
  if (io->n_rows < 0 || io->n_steps < 1 || !io->obs) return MACJD_ERR_INVALID_ARG;
  if (w->obs_dim < 1 || w->obs_pad < w->obs_dim || w->obs_pad % kKC) return MACJD_ERR_INVALID_ARG;
  if (w->hidden < 64 || w->hidden % 64 || w->hidden > 256) return MACJD_ERR_UNSUPPORTED;
  if (w->actor_hidden < 64 || w->actor_hidden % 64 || w->actor_hidden > 256) return MACJD_ERR_UNSUPPORTED;
  if (w->n_actions < 1 || w->n_actions > 64) return MACJD_ERR_UNSUPPORTED;
  if (!w->wa1t || !w->wa2t || !w->wa3t || !w->wfc1t || !w->wrzt || !w->wint || !w->whnt || !w->wqt || !w->w1a ||
      !w->w1p || !w->w2 || !w->bq2)
    return MACJD_ERR_INVALID_ARG;
  if (io->actions && !io->test_mode && !(io->epsilon >= 0.f)) return MACJD_ERR_INVALID_ARG;
  if (io->n_rows == 0) return MACJD_OK;
  AgentArgs a;
  a.w = *w;
  a.io = *io;
  const size_t limit = 200 * 1024;
  // 64-row tiles when they fit in shared memory and still give every SM a CTA
  if (io->tile_rows != 0 && io->tile_rows != 32 && io->tile_rows != 64) return MACJD_ERR_INVALID_ARG;
  bool big = agent_smem_bytes(*w, 64) <= limit && io->n_rows >= 64 * kNumSMs / 2;
  if (io->tile_rows == 64) { if (agent_smem_bytes(*w, 64) > limit) return MACJD_ERR_UNSUPPORTED; big = true; }
  if (io->tile_rows == 32) big = false;
  if (big) {
    const size_t smem = agent_smem_bytes(*w, 64);
    auto k = agent_forward_kernel<4>;
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return MACJD_ERR_CUDA;
    MACJD_LAUNCH(k, (io->n_rows + 63) / 64, kAgentThreads, smem, (cudaStream_t)ctx->stream, a);
  } else {
    const size_t smem = agent_smem_bytes(*w, 32);
    if (smem > limit) return MACJD_ERR_UNSUPPORTED;
    auto k = agent_forward_kernel<2>;
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return MACJD_ERR_CUDA;
    MACJD_LAUNCH(k, (io->n_rows + 31) / 32, kAgentThreads, smem, (cudaStream_t)ctx->stream, a);
  }
  return MACJD_OK;
}

}  // namespace macjd
