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
// Structure: every dense layer is cut into *passes* (one 64-column block of one layer = one
// accumulator lifetime) and every pass into 32-row weight *chunks*.  A single loop walks the
// flat chunk sequence of a timestep: weights stream from the packed K-major copy (L2
// resident) through a 3-stage cp.async ring, one __syncthreads per chunk, and the loop body
// -- a register-tiled 4x4 outer-product microkernel with double-buffered shared-memory
// fragments -- exists exactly once in the binary (instruction-cache friendly).  Pass
// epilogues do the layer-specific work in registers: ReLU + store, the GRU gate algebra, and
// the two small-N heads (actor output layer, per-action Q tail) as in-register partial dot
// products reduced across the 16 column lanes with warp shuffles.
// Activations are row-major in shared memory ([row][K + 4]).
// Bound: FP32 FMA pipe (AI ~ 350 FLOP/B).
#pragma once
#include "macjd_common.cuh"

namespace macjd {

constexpr int kKC = 32;        // K rows per staged weight chunk
constexpr int kNC = 64;        // output columns per pass
constexpr int kStages = 3;     // cp.async ring depth
constexpr int kMaxPasses = 64;

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

enum AgentEpi : int {
  kEpiReluStore = 0,  // out[row][col] = relu(acc + b)              (actor.0, fc1)
  kEpiActorHead,      // a2 = relu(acc + b); Ps += a2 . W_a3        (actor.2 + actor.4)
  kEpiGateR,          // r = sigmoid(acc + b)  -> registers
  kEpiGateZ,          // z = sigmoid(acc + b)  -> registers
  kEpiGateIn,         // keep W_in xf + b_in   -> registers
  kEpiGateHn,         // h' = (1 - z) tanh(in + r (acc + b_hn)) + z h
  kEpiQHead           // pre = acc + b; Qs += per-action tail
};

// Activation source of a pass segment
enum AgentSrc : int { kSrcX = 0, kSrcB0 = 1, kSrcHcur = 2, kSrcHnew = 3 };

struct AgentPass {
  const float* w[2];   // global, K-major, already offset to the pass's first column
  int src[2];          // AgentSrc feeding each segment
  int k[2];            // rows (K) per segment, multiples of 32
  int nseg;
  int ldw;
  int col0;            // first output column / hidden unit of this pass
  int epi;
  int last;            // last pass of its layer (kEpiActorHead: apply the sigmoid)
  const float* bias;
};

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// sum over the 16 column lanes (tx) that share a row group: lanes 0..15 / 16..31 of a warp
__device__ __forceinline__ float reduce16(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  return v;
}

template <int NT>
__global__ void __launch_bounds__(NT, 1) agent_forward_kernel(const AgentArgs a) {
  grid_dependency_wait();   // no-op unless launched as a programmatic dependent (MACJD_LAUNCH)
  constexpr int TM = NT / 4;           // 16 column lanes x (NT/16) row groups x 4 rows
  const macjd_agent_weights& W = a.w;
  const macjd_agent_io& io = a.io;
  const int O = W.obs_dim, Op = W.obs_pad, H = W.hidden, AH = W.actor_hidden, A = W.n_actions;
  const int M = io.n_rows, T = io.n_steps;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int row0 = blockIdx.x * TM;
  const int valid = min(TM, M - row0);
  const int HB = H > AH ? H : AH;
  const int LDX = Op + 4, LDB = HB + 4;

  MACJD_DYNAMIC_SMEM(float, smem);
  float* Xs = smem;                                   // [TM][LDX]  observation tile
  float* B0 = Xs + (size_t)TM * LDX;                  // [TM][LDB]  a1 -> xf
  float* B1 = B0 + (size_t)TM * LDB;                  // [TM][LDB]  h'
  float* B2 = B1 + (size_t)TM * LDB;                  // [TM][LDB]  h (previous step); swaps with B1
  float* Ps = B2 + (size_t)TM * LDB;                  // [A][TM]    actor parameters
  float* Qs = Ps + (size_t)A * TM;                    // [A][TM]    Q per action
  float* wst = Qs + (size_t)A * TM;                   // [3][32][64] weight ring
  AgentPass* passes = reinterpret_cast<AgentPass*>(wst + kStages * kKC * kNC);
  __shared__ int n_passes_s;

  // ---- the pass table of one timestep
  if (tid == 0) {
    int n = 0;
    auto add = [&](const float* w0, int s0, int k0, const float* w1, int s1, int k1, int ldw, int col0, int epi,
                   int last, const float* bias) {
      AgentPass& p = passes[n++];
      p.w[0] = w0; p.src[0] = s0; p.k[0] = k0;
      p.w[1] = w1; p.src[1] = s1; p.k[1] = k1;
      p.nseg = w1 ? 2 : 1; p.ldw = ldw; p.col0 = col0; p.epi = epi; p.last = last; p.bias = bias;
    };
    for (int c = 0; c < AH; c += kNC) add(W.wa1t + c, kSrcX, Op, nullptr, 0, 0, AH, c, kEpiReluStore, 0, W.ba1);
    for (int c = 0; c < AH; c += kNC)
      add(W.wa2t + c, kSrcB0, AH, nullptr, 0, 0, AH, c, kEpiActorHead, c + kNC >= AH, W.ba2);
    for (int c = 0; c < H; c += kNC) add(W.wfc1t + c, kSrcX, Op, nullptr, 0, 0, H, c, kEpiReluStore, 0, W.bfc1);
    for (int u = 0; u < H; u += kNC) {
      const float* whh_rz = W.wrzt + (size_t)H * 2 * H;      // rows H..2H-1: the W_hh part
      add(W.wrzt + u, kSrcB0, H, whh_rz + u, kSrcHcur, H, 2 * H, u, kEpiGateR, 0, W.brz);
      add(W.wrzt + H + u, kSrcB0, H, whh_rz + H + u, kSrcHcur, H, 2 * H, u, kEpiGateZ, 0, W.brz + H);
      add(W.wint + u, kSrcB0, H, nullptr, 0, 0, H, u, kEpiGateIn, 0, W.bin);
      add(W.whnt + u, kSrcHcur, H, nullptr, 0, 0, H, u, kEpiGateHn, 0, W.bhn);
    }
    for (int c = 0; c < H; c += kNC) add(W.wqt + c, kSrcHnew, H, nullptr, 0, 0, H, c, kEpiQHead, 0, W.bq1);
    n_passes_s = n;
  }

  // ---- recurrent state tile (row-major): 16-byte async copies, zero rows past the end
  {
    const int vec = H >> 2;
    const float* hsrc = io.hidden_in ? io.hidden_in : io.hidden;      // initial state: a separate read-only source, or in place
    const bool have = hsrc && !io.hidden_zero_init;
    for (int idx = tid; idx < TM * vec; idx += NT) {
      const int r = idx / vec, k4 = (idx - r * vec) << 2;
      float* dst = B2 + (size_t)r * LDB + k4;
      if (have && r < valid) cp_async16(dst, hsrc + (size_t)(row0 + r) * H + k4);
      else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cp_async_commit();
    cp_async_wait<0>();
  }
  __syncthreads();
  const int n_passes = n_passes_s;

  for (int t = 0; t < T; ++t) {
    const size_t tM = (size_t)t * M;
    float* Hcur = B2;
    float* Hnew = B1;
    // ---- observation tile, zero padded to [TM][Op]; init the two small-head accumulators
    {
      const size_t og = io.obs_group > 1 ? (size_t)io.obs_group : 1;     // rows [k og, (k + 1) og) share observation row k
#pragma unroll 4
      for (int idx = tid; idx < TM * Op; idx += NT) {
        const int r = idx / Op, k = idx - r * Op;
        float v = 0.f;
        if (r < valid && k < O) v = __ldg(io.obs + ((tM + row0 + r) / og) * O + k);
        Xs[(size_t)r * LDX + k] = v;
      }
      const float bq2 = __ldg(W.bq2);
      for (int idx = tid; idx < A * TM; idx += NT) {
        Ps[idx] = __ldg(W.ba3 + idx / TM);
        Qs[idx] = bq2;
      }
    }

    // ---- the chunk pipeline over all passes of this timestep
    float acc[4][4], rg[4][4], zg[4][4], gin[4][4];
    int lp = 0, ls = 0, lc = 0;          // load cursor: pass, segment, chunk
    int cp = 0, cs = 0, cc = 0;          // compute cursor
    int chunk_no = 0;
    auto issue = [&](int stage) {
      if (lp < n_passes) {
        const AgentPass& p = passes[lp];
        const float* src = p.w[ls] + (size_t)(lc * kKC) * p.ldw + tx * 4;
        float* dst = wst + stage * (kKC * kNC) + tx * 4;
#pragma unroll
        for (int rr = ty; rr < kKC; rr += NT / 16) cp_async16(dst + rr * kNC, src + (size_t)rr * p.ldw);
        if (++lc * kKC >= p.k[ls]) { lc = 0; if (++ls >= p.nseg) { ls = 0; ++lp; } }
      }
      cp_async_commit();
    };
    issue(0);
    issue(1);
    while (cp < n_passes) {
      cp_async_wait<1>();
      __syncthreads();
      issue((chunk_no + 2) % kStages);
      const AgentPass& p = passes[cp];
      if (cs == 0 && cc == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      }
      const int src = p.src[cs];
      const float* abase = src == kSrcX ? Xs : src == kSrcB0 ? B0 : src == kSrcHcur ? Hcur : Hnew;
      const int lda = src == kSrcX ? LDX : LDB;
      const float* ab = abase + (size_t)(ty * 4) * lda + cc * kKC;
      const float* wb = wst + (chunk_no % kStages) * (kKC * kNC) + tx * 4;
      // 8 groups of 4 k: fragments double-buffered in registers
      float4 af[2][4], wf[2][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        af[0][i] = *reinterpret_cast<const float4*>(ab + (size_t)i * lda);
        wf[0][i] = *reinterpret_cast<const float4*>(wb + i * kNC);
      }
#pragma unroll
      for (int g = 0; g < kKC / 4; ++g) {
        const int cur = g & 1, nxt = cur ^ 1;
        if (g + 1 < kKC / 4) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            af[nxt][i] = *reinterpret_cast<const float4*>(ab + (size_t)i * lda + (g + 1) * 4);
            wf[nxt][i] = *reinterpret_cast<const float4*>(wb + ((g + 1) * 4 + i) * kNC);
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float av[4] = {af[cur][i].x, af[cur][i].y, af[cur][i].z, af[cur][i].w};
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            acc[i][0] = fmaf(av[kk], wf[cur][kk].x, acc[i][0]);
            acc[i][1] = fmaf(av[kk], wf[cur][kk].y, acc[i][1]);
            acc[i][2] = fmaf(av[kk], wf[cur][kk].z, acc[i][2]);
            acc[i][3] = fmaf(av[kk], wf[cur][kk].w, acc[i][3]);
          }
        }
      }
      ++chunk_no;
      // advance the compute cursor; run the epilogue when the pass is complete
      bool done = false;
      if (++cc * kKC >= p.k[cs]) { cc = 0; if (++cs >= p.nseg) { cs = 0; done = true; } }
      if (done) {
        const int col = p.col0 + tx * 4;
        const float4 b4 = *reinterpret_cast<const float4*>(p.bias + col);
        const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
        switch (p.epi) {
          case kEpiReluStore: {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 v = make_float4(fmaxf(acc[i][0] + bv[0], 0.f), fmaxf(acc[i][1] + bv[1], 0.f),
                                           fmaxf(acc[i][2] + bv[2], 0.f), fmaxf(acc[i][3] + bv[3], 0.f));
              *reinterpret_cast<float4*>(B0 + (size_t)(ty * 4 + i) * LDB + col) = v;
            }
          } break;
          case kEpiActorHead: {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[i][j] = fmaxf(acc[i][j] + bv[j], 0.f);
            for (int act = 0; act < A; ++act) {
              float w3[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) w3[j] = __ldg(W.wa3t + (size_t)(col + j) * A + act);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float s = acc[i][0] * w3[0];
                s = fmaf(acc[i][1], w3[1], s); s = fmaf(acc[i][2], w3[2], s); s = fmaf(acc[i][3], w3[3], s);
                s = reduce16(s);
                if (tx == 0) {
                  float* ps = Ps + act * TM + ty * 4 + i;
                  const float tot = *ps + s;
                  *ps = p.last ? sigmoid_f(tot) : tot;
                }
              }
            }
          } break;
          case kEpiGateR: {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) rg[i][j] = sigmoid_f(acc[i][j] + bv[j]);
          } break;
          case kEpiGateZ: {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) zg[i][j] = sigmoid_f(acc[i][j] + bv[j]);
          } break;
          case kEpiGateIn: {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) gin[i][j] = acc[i][j] + bv[j];
          } break;
          case kEpiGateHn: {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = ty * 4 + i;
              const float4 ho = *reinterpret_cast<const float4*>(Hcur + (size_t)r * LDB + col);
              const float hv[4] = {ho.x, ho.y, ho.z, ho.w};
              float o[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float n = tanhf(gin[i][j] + rg[i][j] * (acc[i][j] + bv[j]));
                o[j] = (1.0f - zg[i][j]) * n + zg[i][j] * hv[j];
              }
              const float4 v = make_float4(o[0], o[1], o[2], o[3]);
              *reinterpret_cast<float4*>(Hnew + (size_t)r * LDB + col) = v;
              if (r < valid) {
                const size_t off = (size_t)(row0 + r) * H + col;
                if (io.hidden_seq) *reinterpret_cast<float4*>(io.hidden_seq + tM * H + off) = v;
                if (io.hidden && t == T - 1) *reinterpret_cast<float4*>(io.hidden + off) = v;
              }
            }
          } break;
          case kEpiQHead: {
            const float4 p4 = *reinterpret_cast<const float4*>(W.w1p + col);
            const float4 w4 = *reinterpret_cast<const float4*>(W.w2 + col);
            const float w1p[4] = {p4.x, p4.y, p4.z, p4.w}, w2[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[i][j] += bv[j];
            for (int act = 0; act < A; ++act) {
              const float4 a4 = *reinterpret_cast<const float4*>(W.w1a + (size_t)act * H + col);
              const float w1a[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float pa = Ps[act * TM + ty * 4 + i];
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) s = fmaf(w2[j], fmaxf(acc[i][j] + w1a[j] + pa * w1p[j], 0.f), s);
                s = reduce16(s);
                if (tx == 0) Qs[act * TM + ty * 4 + i] += s;
              }
            }
          } break;
        }
        ++cp;
      }
    }
    cp_async_wait<0>();
    __syncthreads();

    // ---- outputs: dense [row][action] tables (coalesced), then one thread per row selects
    if (io.q_all || io.params_all) {
      for (int idx = tid; idx < valid * A; idx += NT) {
        const int r = idx / A, act = idx - r * A;
        const size_t o = (tM + row0) * A + idx;
        if (io.q_all) io.q_all[o] = Qs[act * TM + r];
        if (io.params_all) io.params_all[o] = Ps[act * TM + r];
      }
    }
    if (tid < valid) {
      const int r = tid;
      const size_t m = tM + row0 + r;
      const uint8_t* av = io.avail ? io.avail + m * A : nullptr;
      float best = -INFINITY, bestm = -INFINITY;
      int bi = 0, bim = 0, n_avail = 0;
      for (int act = 0; act < A; ++act) {
        const float q = Qs[act * TM + r];
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
        io.q_sel[m] = Qs[s * TM + r];
      }
      if (io.actions) {
        int chosen = bim;
        if (!io.test_mode) {                                   // action_selectors.py:39-57
          const uint32_t row_id = (uint32_t)(io.rng_row_offset + row0 + r);
          const float u = io.u_eps ? io.u_eps[m] : philox_uniform(io.seed, kStreamEpsilon, row_id, (io.rng_step_dev ? __ldg(io.rng_step_dev) : io.rng_step) + t, 0);
          if (u < (io.epsilon_dev ? __ldg(io.epsilon_dev + t) : io.epsilon)) {
            if (io.rand_actions) {
              chosen = io.rand_actions[m];
            } else {
              // uniform over the available actions (multinomial over the 0/1 mask)
              const float u2 = philox_uniform(io.seed, kStreamRandomAction, row_id, (io.rng_step_dev ? __ldg(io.rng_step_dev) : io.rng_step) + t, 0);
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
        const float pw = Ps[chosen * TM + r];
        if (io.power) io.power[m] = pw;
        if (io.actions_mirror) io.actions_mirror[m] = chosen;      // e.g. the caller's page-locked host copy
        if (io.power_mirror) io.power_mirror[m] = pw;
        if (io.q_chosen) io.q_chosen[m] = Qs[chosen * TM + r];
      }
    }
    __syncthreads();
    // h' becomes h for the next unrolled step
    float* tmp = B1; B1 = B2; B2 = tmp;
  }
}

inline size_t agent_smem_bytes(const macjd_agent_weights& w, int TM) {
  const int HB = w.hidden > w.actor_hidden ? w.hidden : w.actor_hidden;
  return sizeof(float) * ((size_t)TM * (w.obs_pad + 4) + (size_t)3 * TM * (HB + 4) + (size_t)2 * w.n_actions * TM +
                          kStages * kKC * kNC) + sizeof(AgentPass) * kMaxPasses;
}

inline int agent_pass_count(const macjd_agent_weights& w) {
  return 2 * (w.actor_hidden / kNC) + 6 * (w.hidden / kNC);
}

inline int agent_launch(const macjd_ctx* ctx, const macjd_agent_weights* w, const macjd_agent_io* io) {
  if (!ctx || !w || !io) return MACJD_ERR_INVALID_ARG;
  if (io->n_rows < 0 || io->n_steps < 1 || !io->obs) return MACJD_ERR_INVALID_ARG;
  if (w->obs_dim < 1 || w->obs_pad < w->obs_dim || w->obs_pad % kKC) return MACJD_ERR_INVALID_ARG;
  if (w->hidden < 64 || w->hidden % 64 || w->hidden > 256) return MACJD_ERR_UNSUPPORTED;
  if (w->actor_hidden < 64 || w->actor_hidden % 64 || w->actor_hidden > 256) return MACJD_ERR_UNSUPPORTED;
  if (w->n_actions < 1 || w->n_actions > 64) return MACJD_ERR_UNSUPPORTED;
  if (agent_pass_count(*w) > kMaxPasses) return MACJD_ERR_UNSUPPORTED;
  if (!w->wa1t || !w->wa2t || !w->wa3t || !w->wfc1t || !w->wrzt || !w->wint || !w->whnt || !w->wqt || !w->w1a ||
      !w->w1p || !w->w2 || !w->bq2 || !w->ba1 || !w->ba2 || !w->ba3 || !w->bfc1 || !w->brz || !w->bin || !w->bhn || !w->bq1)
    return MACJD_ERR_INVALID_ARG;
  if (io->actions && !io->test_mode && !(io->epsilon >= 0.f)) return MACJD_ERR_INVALID_ARG;
  if (io->tile_rows != 0 && io->tile_rows != 8 && io->tile_rows != 16 && io->tile_rows != 32 && io->tile_rows != 64)
    return MACJD_ERR_INVALID_ARG;
  if (io->n_rows == 0) return MACJD_OK;
  AgentArgs a;
  a.w = *w;
  a.io = *io;
  const size_t limit = 200 * 1024;
  // Rows per CTA: 64 (256 threads) when that still gives most SMs a CTA, else 32.  The time of
  // one step of one CTA does not depend on the tile height (every thread always owns a 4x4
  // tile and walks the same chunk sequence), so smaller tiles (16 / 8 rows, selectable with
  // tile_rows) only help by occupying more SMs.
  int tm = io->tile_rows;
  if (tm == 0) tm = io->n_rows >= 64 * kNumSMs / 2 ? 64 : 32;
  while (tm > 8 && agent_smem_bytes(*w, tm) > limit) tm >>= 1;
  const size_t smem = agent_smem_bytes(*w, tm);
  if (smem > limit) return MACJD_ERR_UNSUPPORTED;
  if (io->tile_rows != 0 && tm != io->tile_rows) return MACJD_ERR_UNSUPPORTED;
  const int grid = (io->n_rows + tm - 1) / tm;
  cudaStream_t st = (cudaStream_t)ctx->stream;
#define MACJD_AGENT_LAUNCH(NT)                                                                                       \
  {                                                                                                                  \
    auto k = agent_forward_kernel<NT>;                                                                               \
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)              \
      return MACJD_ERR_CUDA;                                                                                         \
    MACJD_LAUNCH(k, grid, NT, smem, st, a);                                                                          \
  }
  switch (tm) {
    case 64: MACJD_AGENT_LAUNCH(256) break;
    case 32: MACJD_AGENT_LAUNCH(128) break;
    case 16: MACJD_AGENT_LAUNCH(64) break;
    default: MACJD_AGENT_LAUNCH(32) break;
  }
#undef MACJD_AGENT_LAUNCH
  return MACJD_OK;
}

}  // namespace macjd
