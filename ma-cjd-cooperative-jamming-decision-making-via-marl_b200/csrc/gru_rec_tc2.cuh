// The GRU recurrence of a time-unrolled agent pass (core/qmix.py:217-280 -> core/networks.py:88-114) as ONE launch on
// tcgen05 with CTA pairs, for hidden widths H = 128 or 256 (BASELINE config 4: rnn_hidden_dim = 256).
//
// In the learner's unrolls only h_t -> h_t+1 is sequential.  macjd_agent_unroll (agent_unroll.cuh) computes the GRU's
// input products gate_x = relu(fc1 obs) W_i{r,z,n}^T for all T x M rows as dense layers; what is left per timestep
// is  gh = h_{t-1} W_h{r,z,n}^T  ([M][3H], K = H) and the gates.  As one GEMM + one gate kernel per timestep that is
// 2 T launches of ~40 us each at M = 2048 (ncu: profiles/r2s_launches_learner_c4.csv) -- latency, not work.  Here a
// CTA pair keeps 128 rows of h in shared memory (TF32 hi + lo operand tiles, 64 rows per CTA) for the whole unroll,
// streams W_hh (packed 128 x 32 chunks, hi + lo, 1.5 MB at H = 256) from L2 through a TMA ring every step, accumulates
// the three gate products in tensor memory (3 H / 2 columns of the pair's 2-SM layout) and applies the gates in the
// epilogue warps: the recurrence runs at the tensor pipe's rate (576 pair MMAs per step at H = 256) with no launch,
// no global round trip of h and no grid-wide synchronisation -- rows are independent, so pairs never talk.
//
// Per step the MMAs run in NB = H / 128 blocks of 128 output units (r, z, n of the block back to back); the gate
// epilogue of block b overlaps the MMAs of block b + 1.  Only the last block may write h' into the operand tile (the
// earlier blocks' MMAs-in-flight still read h), so earlier blocks hold their 32 values per thread in registers.
// Shared memory at H = 256: 128 KB operand tile + 3 ring stages of 32 KB = 224 KB; gate biases come through L1.
#pragma once
#include "agent_act_tc2.cuh"

#ifndef MACJD_TEST_HOST_EMULATION
namespace macjd {
namespace tc {

struct RecArgs {
  const float* gate_x;      // [T][M][3H] input products (r | z | n) INCLUDING the biases b_ir + b_hr | b_iz + b_hz | b_in
  const float* h0;          // [M][H] or NULL (zeros)
  float* hidden_seq;        // [T][M][H]
  float* hidden_out;        // [M][H] or NULL: h_T
  const float* bhn;         // [H] b_hn (inside the reset gate's product: core/networks.py:88-114 -> nn.GRUCell)
  int M, T;
  alignas(64) CUtensorMap wmap;   // rec_chunks as [bytes / 128][32]
};

template <int NB>
struct RecCfg {
  static constexpr int H = 128 * NB;
  static constexpr int kStages = NB == 1 ? 4 : 3;
  static constexpr int kStagesPerLayer = H / 64;                    // a stage = two consecutive 32-k chunks
  static constexpr int kStagesPerStep = 3 * NB * kStagesPerLayer;
  static constexpr uint32_t kTmemCols = NB == 1 ? 256 : 512;        // 3 NB accumulators of 64 columns
};

template <int NB>
struct RecSmem {
  float hhi[kTcRows * 128 * NB], hlo[kTcRows * 128 * NB];
  unsigned char wst[RecCfg<NB>::kStages][kT2StageBytes];
  uint64_t w_full[4], w_empty[4];
  uint64_t a_ready, d_ready[2];
  uint32_t tmem_base;
  alignas(16) float bhn[128 * NB];
};

// number of 32 KB chunks (hi + lo) of the packed recurrent weights: per 128-unit block, per gate, H / 32 chunks
inline size_t gru_rec_chunk_bytes(int H) { return (size_t)3 * (H / 128) * (H / 32) * kTcChunkBytes; }

template <int NB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kT2Threads, 1) gru_recurrence_tc2_kernel(const __grid_constant__ RecArgs p) {
  using Cfg = RecCfg<NB>;
  constexpr int H = Cfg::H;
  extern __shared__ __align__(1024) unsigned char rec_raw[];
  RecSmem<NB>& S = *reinterpret_cast<RecSmem<NB>*>(rec_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int M = p.M, T = p.T;
  const int row0 = blockIdx.x * kTcRows;
  const int valid = max(0, min(kTcRows, M - row0));

  if (tid == 0) {
    for (int s = 0; s < 4; ++s) { mbar_init(&S.w_full[s], 1); mbar_init(&S.w_empty[s], 1); }
    mbar_init(&S.a_ready, 2 * kT2EpiThreads);
    mbar_init(&S.d_ready[0], 1);
    mbar_init(&S.d_ready[1], 1);
    fence_mbar_init();
  }
  if (warp == kT2EpiWarps + 1) tmem_alloc_2sm(&S.tmem_base, Cfg::kTmemCols);
  for (int i = tid; i < H; i += kT2Threads) S.bhn[i] = __ldg(p.bhn + i);
  fence_before_sync();
  cluster_sync_all();
  fence_after_sync();
  const uint32_t tmem = S.tmem_base;

  if (warp == kT2EpiWarps) {
    // =========================================================== weight stream: this CTA's half of every chunk
    if (lane == 0) {
      uint32_t empty_par = 0;
      int s = 0;
      for (int t = 0; t < T; ++t) {
        for (int L = 0; L < Cfg::kStagesPerStep; ++L) {
          if (t > 0 || L >= Cfg::kStages) { mbar_wait_cluster(&S.w_empty[s], (empty_par >> s) & 1u); empty_par ^= 1u << s; }
          const int row = (int)(((size_t)L * 2 * kTcChunkBytes + (size_t)rank * kT2HalfBytes) / 128);
          if (rank == 0) mbar_expect_tx(&S.w_full[s], 2 * kT2StageBytes);       // both CTAs' halves
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            tma_half_chunk_2sm(S.wst[s] + sub * kT2SubBytes, &p.wmap, row + sub * (kTcChunkBytes / 128), &S.w_full[s], 0);               // hi
            tma_half_chunk_2sm(S.wst[s] + sub * kT2SubBytes + kT2HalfBytes, &p.wmap, row + sub * (kTcChunkBytes / 128) + kTcChunkBytes / 256,
                               &S.w_full[s], 0);                                                                                        // lo
          }
          s = (s + 1 == Cfg::kStages) ? 0 : s + 1;
        }
      }
    }
  } else if (warp == kT2EpiWarps + 1) {
    if (rank == 0) {
      // =========================================================== leader: MMA issue for the pair (warp-uniform, one elected lane)
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
      const uint32_t idesc = umma_idesc_tf32(2 * kTcRows, 128);
      const uint32_t hh = smem_u32(S.hhi), hl = smem_u32(S.hlo), wb0 = smem_u32(S.wst[0]);
      uint32_t full_par = 0, a_par = 0;
      int s = 0;
      for (int t = 0; t < T; ++t) {
        mbar_wait_cluster(&S.a_ready, a_par);          // h_{t-1} is in the operand tiles of both CTAs
        a_par ^= 1u;
        if (lane == 0) TC_STAMP(32);
        for (int L = 0; L < Cfg::kStagesPerStep; ++L) {
          const uint32_t layer = (uint32_t)L / Cfg::kStagesPerLayer, hf = (uint32_t)L % Cfg::kStagesPerLayer;   // layer = 3 block + gate
          mbar_wait_cluster(&S.w_full[s], (full_par >> s) & 1u);
          full_par ^= 1u << s;
          fence_after_sync();
          const uint32_t wbase = wb0 + (uint32_t)s * kT2StageBytes;
          const uint32_t d = tmem_u + 64u * layer;
          if (elect_one()) {
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
              const uint32_t koff = (2 * hf + sub) * kTcAStep;
              const uint64_t dah = umma_smem_desc(hh + koff, 128, H * 32);
              const uint64_t dal = umma_smem_desc(hl + koff, 128, H * 32);
              const uint64_t dbh = umma_smem_desc(wbase + sub * kT2SubBytes, 128, kTcKc * 32);
              const uint64_t dbl = umma_smem_desc(wbase + sub * kT2SubBytes + kT2HalfBytes, 128, kTcKc * 32);
#pragma unroll
              for (int ks = 0; ks < kTcKc / 8; ++ks) {
                const uint64_t adv = (uint64_t)((ks * 256) >> 4);
                mma_tf32_ss_2sm(d, dah + adv, dbh + adv, idesc, (hf == 0 && sub == 0 && ks == 0) ? 0u : 1u);
                mma_tf32_ss_2sm(d, dal + adv, dbh + adv, idesc, 1u);
                mma_tf32_ss_2sm(d, dah + adv, dbl + adv, idesc, 1u);
              }
            }
            mma_commit_2sm(&S.w_empty[s]);
            if (hf == Cfg::kStagesPerLayer - 1 && layer % 3 == 2) mma_commit_2sm(&S.d_ready[layer / 3]);   // block done
          }
          __syncwarp();
          if (lane == 0 && hf == Cfg::kStagesPerLayer - 1 && layer % 3 == 2) TC_STAMP(33 + (int)layer / 3);   // block issued
          if (lane == 0 && L == 0) TC_STAMP(36);                                                              // first stage issued
          s = (s + 1 == Cfg::kStages) ? 0 : s + 1;
        }
      }
    }
  } else {
    // =========================================================== epilogue warps (thread = row x 32 units of a 128-unit block)
    const int q4 = warp & 3, ch = warp >> 2, half = q4 >> 1;
    const int r = (q4 & 1) * 32 + lane;
    const int ub = half * 64 + ch * kT2Upt;          // first unit of this thread inside a block
    const bool live = r < valid;
    const uint32_t tl = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(ch * kT2Upt);
    uint32_t d_par = 0;
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
      for (int i = 0; i < kT2Upt / 4; ++i) {
        const float4 q = (live && p.h0) ? __ldg(reinterpret_cast<const float4*>(p.h0 + (size_t)(row0 + r) * H + b * 128 + ub) + i)
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
        const float v[4] = {q.x, q.y, q.z, q.w};
        store_split4(S.hhi, S.hlo, r, b * 128 + ub + 4 * i, H, v);
      }
    fence_async_smem();
    fence_before_sync();
    mbar_arrive_cluster(&S.a_ready, 0);

    for (int t = 0; t < T; ++t) {
      const size_t tM = (size_t)t * M;
      EP_STAMP(0);
      const float* gx_row = p.gate_x + (tM + row0 + (live ? r : 0)) * 3 * H + ub;
      // next step's input products: ask L2 for them a whole step ahead (gate_x is streamed from HBM once)
      if (live && t + 1 < T) {
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
          for (int g = 0; g < 3; ++g) prefetch_l2(gx_row + (size_t)M * 3 * H + g * H + b * 128);
      }
      float held[NB > 1 ? NB - 1 : 1][kT2Upt];
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const float4* gx = reinterpret_cast<const float4*>(gx_row + b * 128);
        const uint32_t tb = tl + 64u * 3u * (uint32_t)b;
        // this block's input products (3 gates x 32 units per thread), all requested before the wait: they arrive from L2
        // (prefetched a step ahead) while the block's MMAs finish
        float4 gq[3][kT2Upt / 4];
#pragma unroll
        for (int g = 0; g < 3; ++g)
#pragma unroll
          for (int i = 0; i < kT2Upt / 4; ++i) gq[g][i] = live ? __ldg(gx + g * (H / 4) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        epi_wait(&S.d_ready[b], d_par, warp);
        fence_after_sync();
        EP_STAMP(1 + 2 * b);
#pragma unroll
        for (int c0 = 0; c0 < kT2Upt; c0 += 8) {
          float vr[8], vz[8], vh[8];
          tmem_ld8_nowait(tb + (uint32_t)c0, vr);
          tmem_ld8_nowait(tb + 64u + (uint32_t)c0, vz);
          tmem_ld8_nowait(tb + 128u + (uint32_t)c0, vh);
          tmem_ld_wait();
          reg_fence(vr); reg_fence(vz); reg_fence(vh);
          const float4 r0 = gq[0][c0 / 4], r1 = gq[0][c0 / 4 + 1], z0 = gq[1][c0 / 4], z1 = gq[1][c0 / 4 + 1], n0 = gq[2][c0 / 4], n1 = gq[2][c0 / 4 + 1];
          const float gr_[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
          const float gz_[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
          const float gn_[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int c = b * 128 + ub + c0 + 4 * q;       // unit index in [0, H)
            const uint32_t off = umma_off_bytes(r, c, H) >> 2;
            const float4 oh = *reinterpret_cast<const float4*>(S.hhi + off);
            const float4 ol = *reinterpret_cast<const float4*>(S.hlo + off);
            const float hold[4] = {oh.x + ol.x, oh.y + ol.y, oh.z + ol.z, oh.w + ol.w};
            const float4 bhv = *reinterpret_cast<const float4*>(S.bhn + c);
            const float bh_[4] = {bhv.x, bhv.y, bhv.z, bhv.w};
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int jj = 4 * q + j;
              const float rg = sigmoid_fast(vr[jj] + gr_[jj]);
              const float zg = sigmoid_fast(vz[jj] + gz_[jj]);
              const float n = tanh_fast(gn_[jj] + rg * (vh[jj] + bh_[j]));
              o[j] = (1.0f - zg) * n + zg * hold[j];
            }
            if (b == NB - 1) {
              store_split4(S.hhi, S.hlo, r, c, H, o);      // every MMA of this step has completed
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) held[b][c0 + 4 * q + j] = o[j];
            }
            if (live) {
              const float4 v4 = make_float4(o[0], o[1], o[2], o[3]);
              const size_t offg = (size_t)(row0 + r) * H + c;
              *reinterpret_cast<float4*>(p.hidden_seq + tM * H + offg) = v4;
              if (p.hidden_out && t == T - 1) *reinterpret_cast<float4*>(p.hidden_out + offg) = v4;
            }
          }
        }
        EP_STAMP(2 + 2 * b);
      }
      d_par ^= 1u;
      // the earlier blocks' h' go into the operand tile now that no MMA reads it any more
#pragma unroll
      for (int b = 0; b + 1 < NB; ++b)
#pragma unroll
        for (int i = 0; i < kT2Upt / 4; ++i) {
          const float v[4] = {held[b][4 * i], held[b][4 * i + 1], held[b][4 * i + 2], held[b][4 * i + 3]};
          store_split4(S.hhi, S.hlo, r, b * 128 + ub + 4 * i, H, v);
        }
      fence_async_smem();
      fence_before_sync();
      if (t + 1 < T) mbar_arrive_cluster(&S.a_ready, 0);
      EP_STAMP(5);
    }
  }
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();                   // the leader's MMAs read the peer's shared memory until here
  if (warp == kT2EpiWarps + 1) tmem_dealloc_2sm(tmem, Cfg::kTmemCols);
}

inline bool gru_rec_supported(int H) { return (H == 128 || H == 256) && kTcKc == 32; }

template <int NB>
inline int gru_rec_launch_nb(const macjd_ctx* ctx, RecArgs& p, const float* chunks) {
  const size_t smem = sizeof(RecSmem<NB>) + 1024;
  static PerDeviceMax opted;
  if (!opted.covers(ctx->device, smem)) {
    if (cudaFuncSetAttribute(gru_recurrence_tc2_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return MACJD_ERR_CUDA;
    opted.record(ctx->device, smem);
  }
  if (watchdog_arm(ctx->device) != MACJD_OK) return MACJD_ERR_CUDA;
  if (!encode_chunk_map(&p.wmap, chunks, gru_rec_chunk_bytes(128 * NB))) return MACJD_ERR_CUDA;
  const int pairs = (p.M + 2 * kTcRows - 1) / (2 * kTcRows);
  gru_recurrence_tc2_kernel<NB><<<2 * pairs, kT2Threads, smem, (cudaStream_t)ctx->stream>>>(p);
  return MACJD_OK;
}

inline int gru_rec_launch(const macjd_ctx* ctx, int H, RecArgs p, const float* chunks) {
  if (!gru_rec_supported(H) || !chunks || p.M < 1 || p.T < 1) return MACJD_ERR_UNSUPPORTED;
  return H == 128 ? gru_rec_launch_nb<1>(ctx, p, chunks) : gru_rec_launch_nb<2>(ctx, p, chunks);
}

}  // namespace tc
}  // namespace macjd
#endif  // !MACJD_TEST_HOST_EMULATION
