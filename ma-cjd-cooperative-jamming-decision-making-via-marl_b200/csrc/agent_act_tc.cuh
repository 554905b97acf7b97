// Fused agent step on the tcgen05 tensor cores (3xTF32, FP32-level accuracy).
//
// Same computation and outputs as agent_forward_kernel (agent_act.cuh) for the reference's
// network width (rnn_hidden_dim = actor_hidden_dim = 128).  Taken because ncu showed the SIMT
// kernel bound by its shared-memory fragment loads at ~27 % of the FP32 pipe
// (profiles/r1a_agent_r1a.summary.csv): the dense layers move to tcgen05.mma, everything
// else stays FP32 SIMT.
//
// One CTA = 64 agent rows, 5 warps:
//   warps 0-3  epilogue: thread (w, l < 16) owns row 16 w + l = TMEM lane 32 w + l.  It reads its
//              accumulator row with tcgen05.ld, applies bias / ReLU / the GRU gate algebra / the
//              actor head / the per-action Q tail entirely in registers, and writes the next
//              layer's A operand (hi and lo TF32 parts) to shared memory in UMMA layout.
//   warp 4     lane 0 streams the packed weight chunks (128 x 16, hi + lo = 16 KB) from L2 with
//              1-D bulk async copies into a 4-stage ring and issues the tcgen05.mma's; mbarriers
//              carry weights-landed / stage-free / accumulator-ready / operand-ready events.
// TMEM (512 columns): actor.0 -> [0,128), fc1 -> [128,256), actor.2 -> [256,384);
// GRU r | z | W_in xf | W_hn h -> [0,512); Q-head pre-activation -> [0,128).
#pragma once
#include "agent_act.cuh"
#include "tc05.cuh"

#ifndef MACJD_TEST_HOST_EMULATION
namespace macjd {
namespace tc {

constexpr int kTcRows = 64;
constexpr int kTcH = 128;            // hidden width this kernel is specialised for
constexpr int kTcKc = 16;            // k per weight chunk
constexpr int kTcStages = 4;
constexpr int kTcChunkFloats = 2 * kTcH * kTcKc;          // hi + lo
constexpr int kTcChunkBytes = kTcChunkFloats * 4;         // 16 KB
constexpr int kTcThreads = 160;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;\n" ::"r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// write 4 consecutive k of one row (hi and lo parts) into a UMMA-layout operand tile
__device__ __forceinline__ void store_split4(float* hi, float* lo, int r, int k, int K, const float (&v)[4]) {
  const uint32_t off = umma_off_bytes(r, k, K) >> 2;
  float4 h, l;
  h.x = tf32_hi(v[0]); h.y = tf32_hi(v[1]); h.z = tf32_hi(v[2]); h.w = tf32_hi(v[3]);
  l.x = v[0] - h.x; l.y = v[1] - h.y; l.z = v[2] - h.z; l.w = v[3] - h.w;
  *reinterpret_cast<float4*>(hi + off) = h;
  *reinterpret_cast<float4*>(lo + off) = l;
}

struct TcSmem {
  float xhi[kTcRows * 32], xlo[kTcRows * 32];           // observation chunk (32 k)
  float b0hi[kTcRows * kTcH], b0lo[kTcRows * kTcH];     // a1 -> xf
  float hhi[kTcRows * kTcH], hlo[kTcRows * kTcH];       // h -> h'
  float wst[kTcStages][kTcChunkFloats];                 // weight ring
  uint64_t w_full[kTcStages], w_empty[kTcStages];
  uint64_t x_full, x_empty, d_ready, a_ready;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kTcThreads, 1) agent_forward_tc_kernel(const AgentArgs a) {
  extern __shared__ __align__(1024) unsigned char tc_raw[];
  TcSmem& S = *reinterpret_cast<TcSmem*>(tc_raw);
  const macjd_agent_weights& W = a.w;
  const macjd_agent_io& io = a.io;
  const int O = W.obs_dim, Op = W.obs_pad, A = W.n_actions;
  constexpr int H = kTcH;
  const int M = io.n_rows, T = io.n_steps;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * kTcRows;
  const int valid = min(kTcRows, M - row0);
  float* Ps = reinterpret_cast<float*>(tc_raw + sizeof(TcSmem));     // [A][64]
  float* Qs = Ps + (size_t)A * kTcRows;                               // [A][64]
  const int nxc = Op / 32;
  const int chunks_per_step = 4 * nxc + 8 + 48 + 8;

  if (warp == 4) tmem_alloc(&S.tmem_base, 512);
  if (tid == 0) {
    for (int s = 0; s < kTcStages; ++s) { mbar_init(&S.w_full[s], 1); mbar_init(&S.w_empty[s], 1); }
    mbar_init(&S.x_full, 128); mbar_init(&S.x_empty, 1); mbar_init(&S.d_ready, 1); mbar_init(&S.a_ready, 128);
    fence_mbar_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = S.tmem_base;

  if (warp == 4) {
    // =========================================================== weight stream + MMA issue
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(kTcRows, H);
      const char* wsrc = reinterpret_cast<const char*>(W.tc_chunks);
      const long long total_chunks = (long long)chunks_per_step * T;
      long long next_load = 0, cur = 0;
      uint32_t full_par = 0, empty_par = 0;       // bit s = parity to wait for on stage s
      uint32_t x_full_par = 0, a_ready_par = 0;
      auto pump = [&]() {                         // keep up to kTcStages chunks in flight
        while (next_load < total_chunks && next_load < cur + kTcStages) {
          const int s = (int)(next_load % kTcStages);
          if (next_load >= kTcStages) { mbar_wait(&S.w_empty[s], (empty_par >> s) & 1u); empty_par ^= 1u << s; }
          mbar_expect_tx(&S.w_full[s], kTcChunkBytes);
          bulk_g2s(S.wst[s], wsrc + (size_t)(next_load % chunks_per_step) * kTcChunkBytes, kTcChunkBytes, &S.w_full[s]);
          ++next_load;
        }
      };
      // one chunk: 2 k-steps x 3 split products into TMEM column `dcol`
      auto mma_chunk = [&](const float* ahi, const float* alo, uint32_t a_sbo, uint32_t a_koff_bytes, uint32_t dcol,
                           bool first) {
        pump();
        const int s = (int)(cur % kTcStages);
        mbar_wait(&S.w_full[s], (full_par >> s) & 1u);
        full_par ^= 1u << s;
        fence_after_sync();
        const uint64_t dah = umma_smem_desc(smem_u32(ahi) + a_koff_bytes, 128, a_sbo);
        const uint64_t dal = umma_smem_desc(smem_u32(alo) + a_koff_bytes, 128, a_sbo);
        const uint64_t dbh = umma_smem_desc(smem_u32(S.wst[s]), 128, kTcKc * 32);
        const uint64_t dbl = umma_smem_desc(smem_u32(S.wst[s]) + kTcH * kTcKc * 4, 128, kTcKc * 32);
#pragma unroll
        for (int ks = 0; ks < kTcKc / 8; ++ks) {
          const uint64_t adv = (uint64_t)((ks * 256) >> 4);
          mma_tf32_ss(tmem + dcol, dah + adv, dbh + adv, idesc, (first && ks == 0) ? 0u : 1u);
          mma_tf32_ss(tmem + dcol, dal + adv, dbh + adv, idesc, 1u);
          mma_tf32_ss(tmem + dcol, dah + adv, dbl + adv, idesc, 1u);
        }
        mma_commit(&S.w_empty[s]);
        ++cur;
      };
      for (int t = 0; t < T; ++t) {
        // actor.0 and fc1 share the observation operand
        for (int xc = 0; xc < nxc; ++xc) {
          mbar_wait(&S.x_full, x_full_par); x_full_par ^= 1u;
          for (int hf = 0; hf < 2; ++hf) {
            mma_chunk(S.xhi, S.xlo, 32 * 32, hf * 512, 0, xc == 0 && hf == 0);
            mma_chunk(S.xhi, S.xlo, 32 * 32, hf * 512, 128, xc == 0 && hf == 0);
          }
          mma_commit(&S.x_empty);
        }
        mma_commit(&S.d_ready);
        // actor.2 on a1
        mbar_wait(&S.a_ready, a_ready_par); a_ready_par ^= 1u;
        for (int kc = 0; kc < 8; ++kc) mma_chunk(S.b0hi, S.b0lo, H * 32, kc * 512, 256, kc == 0);
        mma_commit(&S.d_ready);
        // GRU on xf (B0) and h
        mbar_wait(&S.a_ready, a_ready_par); a_ready_par ^= 1u;
        for (int g = 0; g < 2; ++g) {
          for (int kc = 0; kc < 8; ++kc) mma_chunk(S.b0hi, S.b0lo, H * 32, kc * 512, g * 128, kc == 0);
          for (int kc = 0; kc < 8; ++kc) mma_chunk(S.hhi, S.hlo, H * 32, kc * 512, g * 128, false);
        }
        for (int kc = 0; kc < 8; ++kc) mma_chunk(S.b0hi, S.b0lo, H * 32, kc * 512, 256, kc == 0);
        for (int kc = 0; kc < 8; ++kc) mma_chunk(S.hhi, S.hlo, H * 32, kc * 512, 384, kc == 0);
        mma_commit(&S.d_ready);
        // Q-head on h'
        mbar_wait(&S.a_ready, a_ready_par); a_ready_par ^= 1u;
        for (int kc = 0; kc < 8; ++kc) mma_chunk(S.hhi, S.hlo, H * 32, kc * 512, 0, kc == 0);
        mma_commit(&S.d_ready);
      }
    }
  } else {
    // =========================================================== epilogue warps
    const int r = warp * 16 + lane;                  // row within the tile (lanes < 16)
    const bool has_row = lane < 16;
    const bool live = has_row && r < valid;
    const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);      // this warp's TMEM lane quarter
    uint32_t d_par = 0, x_empty_par = 0;

    // recurrent state -> hi / lo operand tiles
    if (has_row) {
      const bool have = io.hidden && !io.hidden_zero_init && live;
      for (int k = 0; k < H; k += 4) {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (have) {
          const float4 x = *reinterpret_cast<const float4*>(io.hidden + (size_t)(row0 + r) * H + k);
          v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
        }
        store_split4(S.hhi, S.hlo, r, k, H, v);
      }
    }

    for (int t = 0; t < T; ++t) {
      const size_t tM = (size_t)t * M;
      // ---- observation chunks
      for (int xc = 0; xc < nxc; ++xc) {
        if (t > 0 || xc > 0) { mbar_wait(&S.x_empty, x_empty_par); x_empty_par ^= 1u; }
        if (has_row) {
          const float* obs = io.obs + (tM + row0 + r) * O;
          for (int k = 0; k < 32; k += 4) {
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { const int kk = xc * 32 + k + j; v[j] = (live && kk < O) ? __ldg(obs + kk) : 0.f; }
            store_split4(S.xhi, S.xlo, r, k, 32, v);
          }
        }
        fence_async_smem();
        fence_before_sync();
        mbar_arrive(&S.x_full);
      }

      // ---- E1: a1 = relu(D1 + b) -> B0
      mbar_wait(&S.d_ready, d_par); d_par ^= 1u;
      fence_after_sync();
      for (int c0 = 0; c0 < H; c0 += 8) {
        float v[8];
        tmem_ld8(tl + (uint32_t)c0, v);
        if (has_row) {
          float o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = fmaxf(v[j] + __ldg(W.ba1 + c0 + j), 0.f);
          store_split4(S.b0hi, S.b0lo, r, c0, H, o);
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = fmaxf(v[4 + j] + __ldg(W.ba1 + c0 + 4 + j), 0.f);
          store_split4(S.b0hi, S.b0lo, r, c0 + 4, H, o);
        }
      }
      fence_async_smem();
      fence_before_sync();
      mbar_arrive(&S.a_ready);

      // ---- E2: actor head P = sigmoid(relu(D2 + b) W_a3 + b3);  E3: xf = relu(D3 + b) -> B0
      mbar_wait(&S.d_ready, d_par); d_par ^= 1u;
      fence_after_sync();
      for (int a0 = 0; a0 < A; a0 += 8) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = (a0 + j < A) ? __ldg(W.ba3 + a0 + j) : 0.f;
        for (int c0 = 0; c0 < H; c0 += 8) {
          float v[8];
          tmem_ld8(tl + 256u + (uint32_t)c0, v);
#pragma unroll
          for (int n = 0; n < 8; ++n) {
            const float a2 = fmaxf(v[n] + __ldg(W.ba2 + c0 + n), 0.f);
            const float* w3 = W.wa3t + (size_t)(c0 + n) * A + a0;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (a0 + j < A) acc[j] = fmaf(a2, __ldg(w3 + j), acc[j]);
          }
        }
        if (has_row)
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (a0 + j < A) Ps[(a0 + j) * kTcRows + r] = sigmoid_f(acc[j]);
      }
      for (int c0 = 0; c0 < H; c0 += 8) {
        float v[8];
        tmem_ld8(tl + 128u + (uint32_t)c0, v);
        if (has_row) {
          float o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = fmaxf(v[j] + __ldg(W.bfc1 + c0 + j), 0.f);
          store_split4(S.b0hi, S.b0lo, r, c0, H, o);
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = fmaxf(v[4 + j] + __ldg(W.bfc1 + c0 + 4 + j), 0.f);
          store_split4(S.b0hi, S.b0lo, r, c0 + 4, H, o);
        }
      }
      fence_async_smem();
      fence_before_sync();
      mbar_arrive(&S.a_ready);

      // ---- E4: GRU gates -> h' (in place over h), global hidden outputs
      mbar_wait(&S.d_ready, d_par); d_par ^= 1u;
      fence_after_sync();
      for (int c0 = 0; c0 < H; c0 += 8) {
        float vr[8], vz[8], vi[8], vh[8];
        tmem_ld8(tl + (uint32_t)c0, vr);
        tmem_ld8(tl + 128u + (uint32_t)c0, vz);
        tmem_ld8(tl + 256u + (uint32_t)c0, vi);
        tmem_ld8(tl + 384u + (uint32_t)c0, vh);
        if (has_row) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int c = c0 + 4 * q;
            const uint32_t off = umma_off_bytes(r, c, H) >> 2;
            const float4 hh = *reinterpret_cast<const float4*>(S.hhi + off);
            const float4 hl = *reinterpret_cast<const float4*>(S.hlo + off);
            const float hold[4] = {hh.x + hl.x, hh.y + hl.y, hh.z + hl.z, hh.w + hl.w};
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int u = c + j, jj = 4 * q + j;
              const float rg = sigmoid_f(vr[jj] + __ldg(W.brz + u));
              const float zg = sigmoid_f(vz[jj] + __ldg(W.brz + H + u));
              const float n = tanhf(vi[jj] + __ldg(W.bin + u) + rg * (vh[jj] + __ldg(W.bhn + u)));
              o[j] = (1.0f - zg) * n + zg * hold[j];
            }
            store_split4(S.hhi, S.hlo, r, c, H, o);
            if (live) {
              const float4 v4 = make_float4(o[0], o[1], o[2], o[3]);
              const size_t offg = (size_t)(row0 + r) * H + c;
              if (io.hidden_seq) *reinterpret_cast<float4*>(io.hidden_seq + tM * H + offg) = v4;
              if (io.hidden && t == T - 1) *reinterpret_cast<float4*>(io.hidden + offg) = v4;
            }
          }
        }
      }
      fence_async_smem();
      fence_before_sync();
      mbar_arrive(&S.a_ready);

      // ---- E5: Q tail, outputs, selection
      mbar_wait(&S.d_ready, d_par); d_par ^= 1u;
      fence_after_sync();
      const float bq2 = __ldg(W.bq2);
      for (int a0 = 0; a0 < A; a0 += 8) {
        float acc[8], pa[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[j] = bq2; pa[j] = (has_row && a0 + j < A) ? Ps[(a0 + j) * kTcRows + r] : 0.f; }
        for (int c0 = 0; c0 < H; c0 += 8) {
          float v[8];
          tmem_ld8(tl + (uint32_t)c0, v);
#pragma unroll
          for (int n = 0; n < 8; ++n) {
            const int u = c0 + n;
            const float pre = v[n] + __ldg(W.bq1 + u);
            const float w1p = __ldg(W.w1p + u), w2 = __ldg(W.w2 + u);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (a0 + j < A)
                acc[j] = fmaf(w2, fmaxf(pre + __ldg(W.w1a + (size_t)(a0 + j) * H + u) + pa[j] * w1p, 0.f), acc[j]);
          }
        }
        if (has_row)
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (a0 + j < A) Qs[(a0 + j) * kTcRows + r] = acc[j];
      }
      fence_before_sync();          // TMEM reads of this step are complete before the next x_full arrival
      if (live) {
        const size_t m = tM + row0 + r;
        const uint8_t* av = io.avail ? io.avail + m * A : nullptr;
        float best = -INFINITY, bestm = -INFINITY;
        int bi = 0, bim = 0, n_avail = 0;
        for (int act = 0; act < A; ++act) {
          const float q = Qs[act * kTcRows + r];
          const float p = Ps[act * kTcRows + r];
          if (io.q_all) io.q_all[m * A + act] = q;
          if (io.params_all) io.params_all[m * A + act] = p;
          if (q > best) { best = q; bi = act; }
          const bool ok = av ? (av[act] != 0) : true;
          n_avail += ok ? 1 : 0;
          const float qm = ok ? q : -INFINITY;
          if (qm > bestm) { bestm = qm; bim = act; }
        }
        if (io.greedy) io.greedy[m] = bi;
        if (io.sel_actions && io.q_sel) {
          int s = io.sel_actions[m];
          s = s < 0 ? 0 : (s >= A ? A - 1 : s);
          io.q_sel[m] = Qs[s * kTcRows + r];
        }
        if (io.actions) {
          int chosen = bim;
          if (!io.test_mode) {
            const uint32_t row_id = (uint32_t)(row0 + r);
            const float u = io.u_eps ? io.u_eps[m] : philox_uniform(io.seed, kStreamEpsilon, row_id, io.rng_step + t, 0);
            if (u < io.epsilon) {
              if (io.rand_actions) {
                chosen = io.rand_actions[m];
              } else {
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
          if (io.power) io.power[m] = Ps[chosen * kTcRows + r];
          if (io.q_chosen) io.q_chosen[m] = Qs[chosen * kTcRows + r];
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 512);
}

inline size_t agent_tc_smem_bytes(const macjd_agent_weights& w) {
  return sizeof(TcSmem) + sizeof(float) * 2 * (size_t)w.n_actions * kTcRows + 1024;
}

inline bool agent_tc_supported(const macjd_agent_weights& w) {
  return w.tc_chunks != nullptr && w.hidden == kTcH && w.actor_hidden == kTcH && w.obs_pad % 32 == 0 &&
         agent_tc_smem_bytes(w) <= 227 * 1024;
}

inline int agent_tc_launch(const macjd_ctx* ctx, const AgentArgs& a) {
  const size_t smem = agent_tc_smem_bytes(a.w);
  if (cudaFuncSetAttribute(agent_forward_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return MACJD_ERR_CUDA;
  const int grid = (a.io.n_rows + kTcRows - 1) / kTcRows;
  agent_forward_tc_kernel<<<grid, kTcThreads, smem, (cudaStream_t)ctx->stream>>>(a);
  return MACJD_OK;
}

}  // namespace tc
}  // namespace macjd
#endif  // !MACJD_TEST_HOST_EMULATION
