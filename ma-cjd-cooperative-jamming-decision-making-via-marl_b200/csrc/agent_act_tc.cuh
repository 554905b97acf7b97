// Fused agent step on the tcgen05 tensor cores (3xTF32, FP32-level accuracy).
//
// Same computation and outputs as agent_forward_kernel (agent_act.cuh) for the reference's
// network width (rnn_hidden_dim = actor_hidden_dim = 128).  Taken because ncu showed the SIMT
// kernel bound by its shared-memory fragment loads at ~27 % of the FP32 pipe
// (profiles/r1a_agent_r1a.summary.csv): the dense layers move to tcgen05.mma, everything
// else stays FP32 SIMT.
//
// One CTA = 64 agent rows, 6 warps:
//   warps 0-3  epilogue.  Accumulators are M = 64 tiles in TMEM (row i of quarter w on lane
//              32 w + i, i < 16); the warps read them with tcgen05.ld.16x256b, which spreads the 16
//              valid lanes over all 32 threads like an m16n8 accumulator fragment (thread t: rows
//              t/4 and t/4 + 8, column pairs 2 (t%4) of every 8-column group).  Bias / ReLU / GRU gate
//              algebra / actor head / per-action Q tail run in registers (row sums finish with two
//              shuffles over the 4 threads of a row); the next layer's A operand is written to
//              shared memory as TF32 hi + lo tiles in UMMA K-major layout.
//   warp 4     lane 0 streams the packed weight chunks (128 x 32, hi + lo = 32 KB) from L2 with
//              1-D bulk async copies (cp.async.bulk + mbarrier complete_tx) into a 2-stage ring.
//   warp 5     lane 0 issues the tcgen05.mma's: M = 64, N = 128, K = 8, kind::tf32, three split
//              products per k-step, accumulators in 512 TMEM columns.
//   mbarriers carry weights-landed / stage-free / accumulator-ready / operand-ready events.
// Measured on B200 (tools/tc_mma_rate.py): one cta_group::1 M x N x 8 TF32 SS MMA costs 68 cycles at
// N = 128 for M = 64 and for M = 128 -- a 64-row CTA wastes half of every instruction -- and a single
// thread that both waits on barriers and issues copies needs ~800 cycles per chunk, hence the separate
// stream warp.  One CTA-step takes ~32 us.
//
// This is the first tensor-core version (io->path = 2).  The default tensor-core kernel is the
// CTA-pair one in agent_act_tc2.cuh (path 3 / auto): cta_group::2 MMAs (M = 128 per pair at 37 cycles),
// half the weight bytes per CTA, warp-uniform issue, 22 us per step even for <= 64 rows.  This file
// also holds what both share: chunk geometry, TcConst, the operand-tile helpers.
#pragma once
#include "agent_act.cuh"
#include "tc05.cuh"

#ifndef MACJD_TEST_HOST_EMULATION
namespace macjd {
namespace tc {

constexpr int kTcRows = 64;
constexpr int kTcH = 128;            // hidden width this kernel is specialised for
#ifndef MACJD_TC_KC
#define MACJD_TC_KC 32
#endif
constexpr int kTcKc = MACJD_TC_KC;   // k per weight chunk (16 or 32); the ring always holds 64 KB
constexpr int kTcStages = 64 / kTcKc;
constexpr int kTcChunksPerH = kTcH / kTcKc;               // chunks of a K = 128 layer
constexpr int kTcChunksPerX = 32 / kTcKc;                 // chunks per 32-wide observation block
constexpr uint32_t kTcAStep = kTcKc * 32;                 // A-tile byte offset between chunks
constexpr int kTcChunkFloats = 2 * kTcH * kTcKc;          // hi + lo
constexpr int kTcChunkBytes = kTcChunkFloats * 4;         // 16 KB
constexpr int kTcThreads = 192;       // 4 epilogue warps + weight-stream warp + MMA warp

__device__ float g_tc_sink;

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

// Pull one step's weight chunks (about 1 MB, read by every CTA) into L2 before the rings ask for them:
// after an L2 flush every ring stage would otherwise wait a full HBM round trip.
#ifndef MACJD_TC_PREFETCH
#define MACJD_TC_PREFETCH 1
#endif
__device__ __forceinline__ void warm_weights_l2(const float* chunks, int chunks_per_step, int threads) {
  const char* base = reinterpret_cast<const char*>(chunks);
#if MACJD_TC_PREFETCH == 1
  for (size_t line = (size_t)blockIdx.x * threads + threadIdx.x; line < (size_t)chunks_per_step * kTcChunkBytes / 128;
       line += (size_t)gridDim.x * threads)
    prefetch_l2(base + line * 128);
#elif MACJD_TC_PREFETCH == 2
  if (threadIdx.x == 0)
    for (int c = blockIdx.x; c < chunks_per_step; c += gridDim.x)
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(base + (size_t)c * kTcChunkBytes), "r"(kTcChunkBytes) : "memory");
#elif MACJD_TC_PREFETCH == 3
  float acc = 0.f;
  for (size_t line = (size_t)blockIdx.x * threads + threadIdx.x; line < (size_t)chunks_per_step * kTcChunkBytes / 128;
       line += (size_t)gridDim.x * threads)
    acc += *reinterpret_cast<const volatile float*>(base + line * 128);
  if (acc == 1.2345e-33f) g_tc_sink = acc;
#elif MACJD_TC_PREFETCH == 4
  // every line is requested by gridDim.x / 16 CTAs spread over the chip (both L2 partitions)
  const size_t n_lines = (size_t)chunks_per_step * kTcChunkBytes / 128;
  for (size_t i = threadIdx.x; i * 16 + (blockIdx.x & 15) < n_lines; i += threads) prefetch_l2(base + (i * 16 + (blockIdx.x & 15)) * 128);
#elif MACJD_TC_PREFETCH == 5
  const size_t n_lines = (size_t)chunks_per_step * kTcChunkBytes / 128;
  float acc = 0.f;
  for (size_t i = threadIdx.x; i * 16 + (blockIdx.x & 15) < n_lines; i += threads)
    acc += *reinterpret_cast<const volatile float*>(base + (i * 16 + (blockIdx.x & 15)) * 128);
  if (acc == 1.2345e-33f) g_tc_sink = acc;
#else
  (void)base; (void)chunks_per_step; (void)threads;
#endif
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

// Optional phase timestamps of CTA 0 (compile with -DMACJD_TC_PROFILE; tools/tc_phase_profile.py)
#ifdef MACJD_TC_PROFILE
__device__ unsigned long long g_tc_prof[64 + 1024];     // [64 + 2 b], [65 + 2 b]: entry / exit time (ns) of CTA b < 512
__device__ __forceinline__ unsigned long long tc_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TC_CTA_STAMP(which) do { if (threadIdx.x == 0 && blockIdx.x < 512) g_tc_prof[64 + 2 * blockIdx.x + (which)] = tc_globaltimer(); } while (0)
#define TC_STAMP_ONCE(slot) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_tc_prof[slot] = clock64(); } while (0)
#define TC_STAMP(slot) do { if (blockIdx.x == 0 && t == (T > 1 ? 1 : 0)) g_tc_prof[slot] = clock64(); } while (0)
#else
#define TC_STAMP(slot) do { } while (0)
#define TC_CTA_STAMP(which) do { } while (0)
#define TC_STAMP_ONCE(slot) do { } while (0)
#endif

// Small per-layer vectors, staged once per CTA, packed so that one 16-byte shared-memory load
// (a warp-wide broadcast) brings everything an epilogue needs for one hidden unit.
struct TcConst {
  float4 gate_b[kTcH];      // (b_r, b_z, b_in, b_hn) per hidden unit
  float4 q_c[kTcH];         // (fc2_q_head.0.bias, W1[:, H+A], fc2_q_head.2.weight, 0) per unit
  float ba1[kTcH], ba2[kTcH], bfc1[kTcH];
  float4 wa3t[kTcH * 2];    // [unit][8 actions]  actor.4.weight^T, zero padded
  float4 w1a[kTcH * 2];     // [unit][8 actions]  fc2_q_head.0.weight[:, H + a]
  float ba3[8];
};

struct TcSmem {
  float xhi[kTcRows * 32], xlo[kTcRows * 32];           // observation chunk (32 k)
  float b0hi[kTcRows * kTcH], b0lo[kTcRows * kTcH];     // a1 -> xf
  float hhi[kTcRows * kTcH], hlo[kTcRows * kTcH];       // h -> h'
  float wst[kTcStages][kTcChunkFloats];                 // weight ring
  TcConst c;
  uint64_t w_full[kTcStages], w_empty[kTcStages];
  uint64_t x_full, x_empty, d_ready, a_ready;
  uint32_t tmem_base;
};

// TMEM columns (M = 64 accumulators: row i of warp-quarter w on lane 32 w + i, i < 16).  The
// epilogue warps read them with tcgen05.ld.16x256b, which spreads the 16 valid lanes over all
// 32 threads like an m16n8 accumulator fragment: thread t owns rows g and g + 8 (g = t / 4)
// and, in every 8-column group, columns 2 (t % 4) and 2 (t % 4) + 1.
constexpr uint32_t kColA1 = 0, kColFc1 = 128, kColA2 = 256;             // actor.0, fc1, actor.2
constexpr uint32_t kColR = 0, kColZ = 128, kColIn = 256, kColHn = 384;   // GRU
constexpr uint32_t kColQ = 0;
constexpr uint32_t kTmemCols = 512;

// write 2 consecutive k of one row (hi and lo parts) into a UMMA-layout operand tile
__device__ __forceinline__ void store_split2(float* hi, float* lo, int r, int k, int K, float v0, float v1) {
  const uint32_t off = umma_off_bytes(r, k, K) >> 2;
  float2 h, l;
  h.x = tf32_hi(v0); h.y = tf32_hi(v1);
  l.x = v0 - h.x; l.y = v1 - h.y;
  *reinterpret_cast<float2*>(hi + off) = h;
  *reinterpret_cast<float2*>(lo + off) = l;
}

__global__ void __launch_bounds__(kTcThreads, 1) agent_forward_tc_kernel(const AgentArgs a) {
  extern __shared__ __align__(1024) unsigned char tc_raw[];
  TcSmem& S = *reinterpret_cast<TcSmem*>(tc_raw);
  const macjd_agent_weights& W = a.w;
  const macjd_agent_io& io = a.io;
  const int O = W.obs_dim, Op = W.obs_pad, A = W.n_actions;
  constexpr int H = kTcH;
  const int M = io.n_rows, T = io.n_steps;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  TC_CTA_STAMP(0);
  TC_STAMP_ONCE(20);
  const int row0 = blockIdx.x * kTcRows;
  const int valid = min(kTcRows, M - row0);
  float* Ps = reinterpret_cast<float*>(tc_raw + sizeof(TcSmem));     // [A][64]
  float* Qs = Ps + (size_t)A * kTcRows;                               // [A][64]
  const int nxc = Op / 32;
  const int chunks_per_step = 2 * kTcChunksPerX * nxc + 8 * kTcChunksPerH;
  warm_weights_l2(W.tc_chunks, chunks_per_step, kTcThreads);

  if (warp == 5) tmem_alloc(&S.tmem_base, kTmemCols);
  if (tid == 0) {
    for (int s = 0; s < kTcStages; ++s) { mbar_init(&S.w_full[s], 1); mbar_init(&S.w_empty[s], 1); }
    mbar_init(&S.x_full, 128); mbar_init(&S.x_empty, 1); mbar_init(&S.d_ready, 1); mbar_init(&S.a_ready, 128);
    fence_mbar_init();
  }
  // stage the small vectors
  for (int i = tid; i < H; i += kTcThreads) {
    // all loads first (read-only path), then the stores: interleaved, every load waited for the
    // previous shared-memory store (possible aliasing) -- ten serial DRAM round trips after an L2 flush
    const float g0 = __ldg(W.brz + i), g1 = __ldg(W.brz + H + i), g2 = __ldg(W.bin + i), g3 = __ldg(W.bhn + i);
    const float q0 = __ldg(W.bq1 + i), q1 = __ldg(W.w1p + i), q2 = __ldg(W.w2 + i);
    const float b1 = __ldg(W.ba1 + i), b2 = __ldg(W.ba2 + i), b3 = __ldg(W.bfc1 + i);
    float w3[8], wq[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      w3[j] = j < A ? __ldg(W.wa3t + (size_t)i * A + j) : 0.f;
      wq[j] = j < A ? __ldg(W.w1a + (size_t)j * H + i) : 0.f;
    }
    S.c.gate_b[i] = make_float4(g0, g1, g2, g3);
    S.c.q_c[i] = make_float4(q0, q1, q2, 0.f);
    S.c.ba1[i] = b1; S.c.ba2[i] = b2; S.c.bfc1[i] = b3;
    S.c.wa3t[2 * i] = make_float4(w3[0], w3[1], w3[2], w3[3]);
    S.c.wa3t[2 * i + 1] = make_float4(w3[4], w3[5], w3[6], w3[7]);
    S.c.w1a[2 * i] = make_float4(wq[0], wq[1], wq[2], wq[3]);
    S.c.w1a[2 * i + 1] = make_float4(wq[4], wq[5], wq[6], wq[7]);
  }
  if (tid < 8) S.c.ba3[tid] = tid < A ? W.ba3[tid] : 0.f;
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = S.tmem_base;

  if (warp == 4) {
    // =========================================================== weight stream (bulk async copies)
    if (lane == 0) {
      const char* wsrc = reinterpret_cast<const char*>(W.tc_chunks);
      const long long total_chunks = (long long)chunks_per_step * T;
      uint32_t empty_par = 0;                     // bit s = parity to wait for on stage s
      for (long long L = 0; L < total_chunks; ++L) {
        const int s = (int)(L % kTcStages);
        if (L >= kTcStages) { mbar_wait(&S.w_empty[s], (empty_par >> s) & 1u); empty_par ^= 1u << s; }
        mbar_expect_tx(&S.w_full[s], kTcChunkBytes);
        bulk_g2s(S.wst[s], wsrc + (size_t)(L % chunks_per_step) * kTcChunkBytes, kTcChunkBytes, &S.w_full[s]);
      }
    }
  } else if (warp == 5) {
    // =========================================================== MMA issue
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(kTcRows, kTcH);
      long long cur = 0;
      uint32_t full_par = 0;
      uint32_t x_full_par = 0, a_ready_par = 0;
      // one chunk: kTcKc / 8 k-steps x 3 split products (M = 64, N = 128) into TMEM column `dcol`
      auto mma_chunk = [&](const float* ahi, const float* alo, uint32_t a_sbo, uint32_t a_koff_bytes, uint32_t dcol,
                           bool first) {
        const int s = (int)(cur % kTcStages);
        mbar_wait(&S.w_full[s], (full_par >> s) & 1u);
        full_par ^= 1u << s;
        fence_after_sync();
        const uint64_t dah = umma_smem_desc(smem_u32(ahi) + a_koff_bytes, 128, a_sbo);
        const uint64_t dal = umma_smem_desc(smem_u32(alo) + a_koff_bytes, 128, a_sbo);
        const uint32_t wbase = smem_u32(S.wst[s]);
        const uint64_t dbh = umma_smem_desc(wbase, 128, kTcKc * 32);
        const uint64_t dbl = umma_smem_desc(wbase + kTcH * kTcKc * 4, 128, kTcKc * 32);
        const uint32_t d = tmem + dcol;
#pragma unroll
        for (int ks = 0; ks < kTcKc / 8; ++ks) {
          const uint64_t adv = (uint64_t)((ks * 256) >> 4);
          mma_tf32_ss(d, dah + adv, dbh + adv, idesc, (first && ks == 0) ? 0u : 1u);
          mma_tf32_ss(d, dal + adv, dbh + adv, idesc, 1u);
          mma_tf32_ss(d, dah + adv, dbl + adv, idesc, 1u);
        }
        mma_commit(&S.w_empty[s]);
        ++cur;
      };
      for (int t = 0; t < T; ++t) {
        // actor.0 and fc1 share the observation operand
        for (int xc = 0; xc < nxc; ++xc) {
          mbar_wait(&S.x_full, x_full_par); x_full_par ^= 1u;
          TC_STAMP(32);
          for (int hf = 0; hf < kTcChunksPerX; ++hf) {
            mma_chunk(S.xhi, S.xlo, 32 * 32, hf * kTcAStep, kColA1, xc == 0 && hf == 0);
            mma_chunk(S.xhi, S.xlo, 32 * 32, hf * kTcAStep, kColFc1, xc == 0 && hf == 0);
          }
          mma_commit(&S.x_empty);
        }
        mma_commit(&S.d_ready);
        TC_STAMP(33);
        // actor.2 on a1
        mbar_wait(&S.a_ready, a_ready_par); a_ready_par ^= 1u;
        TC_STAMP(34);
        for (int kc = 0; kc < kTcChunksPerH; ++kc) mma_chunk(S.b0hi, S.b0lo, H * 32, kc * kTcAStep, kColA2, kc == 0);
        mma_commit(&S.d_ready);
        TC_STAMP(35);
        // GRU on xf (B0) and h
        mbar_wait(&S.a_ready, a_ready_par); a_ready_par ^= 1u;
        TC_STAMP(36);
        for (int g = 0; g < 2; ++g) {
          for (int kc = 0; kc < kTcChunksPerH; ++kc) mma_chunk(S.b0hi, S.b0lo, H * 32, kc * kTcAStep, g == 0 ? kColR : kColZ, kc == 0);
          for (int kc = 0; kc < kTcChunksPerH; ++kc) mma_chunk(S.hhi, S.hlo, H * 32, kc * kTcAStep, g == 0 ? kColR : kColZ, false);
        }
        for (int kc = 0; kc < kTcChunksPerH; ++kc) mma_chunk(S.b0hi, S.b0lo, H * 32, kc * kTcAStep, kColIn, kc == 0);
        for (int kc = 0; kc < kTcChunksPerH; ++kc) mma_chunk(S.hhi, S.hlo, H * 32, kc * kTcAStep, kColHn, kc == 0);
        mma_commit(&S.d_ready);
        TC_STAMP(37);
        // Q-head on h'
        mbar_wait(&S.a_ready, a_ready_par); a_ready_par ^= 1u;
        TC_STAMP(38);
        for (int kc = 0; kc < kTcChunksPerH; ++kc) mma_chunk(S.hhi, S.hlo, H * 32, kc * kTcAStep, kColQ, kc == 0);
        mma_commit(&S.d_ready);
        TC_STAMP(39);
      }
    }
  } else {
    // =========================================================== epilogue warps
    const int half = lane >> 4;                      // which 64 output units of the row
    const int r = warp * 16 + (lane & 15);           // row within the tile
    const int ub = half * 64;                        // first unit owned by this thread
    const bool live = r < valid;
    const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);      // this warp's TMEM lane quarter
    const int fr = warp * 16 + (lane >> 2);          // fragment rows of this thread: fr and fr + 8
    const int fq = lane & 3;                         // fragment column pair within every 8-column group
    uint32_t d_par = 0, x_empty_par = 0;

    // recurrent state -> hi / lo operand tiles (each thread: its half of the row)
    {
      const float* hsrc = io.hidden_in ? io.hidden_in : io.hidden;
      const bool have = hsrc && !io.hidden_zero_init && live;
      for (int k = ub; k < ub + 64; k += 4) {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (have) {
          const float4 x = *reinterpret_cast<const float4*>(hsrc + (size_t)(row0 + r) * H + k);
          v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
        }
        store_split4(S.hhi, S.hlo, r, k, H, v);
      }
    }

    for (int t = 0; t < T; ++t) {
      const size_t tM = (size_t)t * M;
#ifdef MACJD_TC_PROFILE
#define EP_STAMP(slot) do { if (tid == 0) TC_STAMP(slot); } while (0)
#else
#define EP_STAMP(slot) do { } while (0)
#endif
      EP_STAMP(0);
      // ---- observation chunks (32 k): 16 k per half-warp thread
      for (int xc = 0; xc < nxc; ++xc) {
        if (t > 0 || xc > 0) { mbar_wait(&S.x_empty, x_empty_par); x_empty_par ^= 1u; }
        const float* obs = io.obs + (tM + row0 + r) * O;
        for (int k = half * 16; k < half * 16 + 16; k += 4) {
          float v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) { const int kk = xc * 32 + k + j; v[j] = (live && kk < O) ? __ldg(obs + kk) : 0.f; }
          store_split4(S.xhi, S.xlo, r, k, 32, v);
        }
        fence_async_smem();
        fence_before_sync();
        mbar_arrive(&S.x_full);
      }

      EP_STAMP(1);
      // ---- E1: a1 = relu(D1 + b) -> B0
      mbar_wait(&S.d_ready, d_par); d_par ^= 1u;
      fence_after_sync();
      EP_STAMP(2);
      for (int cb = 0; cb < 4; ++cb) {
        float v[16];
        tmem_ld_16x256b_x4_nowait(tl + kColA1 + (uint32_t)(32 * cb), v);
        tmem_ld_wait();
        reg_fence(v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = 32 * cb + 8 * i + 2 * fq;
          const float2 b = *reinterpret_cast<const float2*>(&S.c.ba1[c]);
#pragma unroll
          for (int hr = 0; hr < 2; ++hr)
            store_split2(S.b0hi, S.b0lo, fr + 8 * hr, c, H, fmaxf(v[4 * i + 2 * hr] + b.x, 0.f),
                         fmaxf(v[4 * i + 2 * hr + 1] + b.y, 0.f));
        }
      }
      fence_async_smem();
      fence_before_sync();
      mbar_arrive(&S.a_ready);

      EP_STAMP(3);
      // ---- E2: actor head P = sigmoid(relu(D2 + b) W_a3 + b3);  E3: xf = relu(D3 + b) -> B0
      mbar_wait(&S.d_ready, d_par); d_par ^= 1u;
      fence_after_sync();
      EP_STAMP(4);
      {
        float acc[2][8];
#pragma unroll
        for (int hr = 0; hr < 2; ++hr)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[hr][j] = 0.f;
        for (int cb = 0; cb < 4; ++cb) {
          float v[16];
          tmem_ld_16x256b_x4_nowait(tl + kColA2 + (uint32_t)(32 * cb), v);
          tmem_ld_wait();
          reg_fence(v);
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int u = 32 * cb + 8 * i + 2 * fq + e;
              const float b = S.c.ba2[u];
              const float4 wl = S.c.wa3t[2 * u], wh = S.c.wa3t[2 * u + 1];
              const float w[8] = {wl.x, wl.y, wl.z, wl.w, wh.x, wh.y, wh.z, wh.w};
#pragma unroll
              for (int hr = 0; hr < 2; ++hr) {
                const float a2 = fmaxf(v[4 * i + 2 * hr + e] + b, 0.f);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[hr][j] = fmaf(a2, w[j], acc[hr][j]);
              }
            }
        }
#pragma unroll
        for (int hr = 0; hr < 2; ++hr)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float tot = acc[hr][j];
            tot += __shfl_xor_sync(0xffffffffu, tot, 1);
            tot += __shfl_xor_sync(0xffffffffu, tot, 2);
            if (fq == 0 && j < A) Ps[j * kTcRows + fr + 8 * hr] = sigmoid_fast(tot + S.c.ba3[j]);
          }
      }
      EP_STAMP(5);
      for (int cb = 0; cb < 4; ++cb) {
        float v[16];
        tmem_ld_16x256b_x4_nowait(tl + kColFc1 + (uint32_t)(32 * cb), v);
        tmem_ld_wait();
        reg_fence(v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = 32 * cb + 8 * i + 2 * fq;
          const float2 b = *reinterpret_cast<const float2*>(&S.c.bfc1[c]);
#pragma unroll
          for (int hr = 0; hr < 2; ++hr)
            store_split2(S.b0hi, S.b0lo, fr + 8 * hr, c, H, fmaxf(v[4 * i + 2 * hr] + b.x, 0.f),
                         fmaxf(v[4 * i + 2 * hr + 1] + b.y, 0.f));
        }
      }
      fence_async_smem();
      fence_before_sync();
      mbar_arrive(&S.a_ready);

      EP_STAMP(6);
      // ---- E4: GRU gates -> h' (in place over h), global hidden outputs
      mbar_wait(&S.d_ready, d_par); d_par ^= 1u;
      fence_after_sync();
      EP_STAMP(7);
      for (int c0 = 0; c0 < H; c0 += 16) {
        float vr[8], vz[8], vi[8], vh[8];
        tmem_ld_16x256b_x2_nowait(tl + kColR + (uint32_t)c0, vr);
        tmem_ld_16x256b_x2_nowait(tl + kColZ + (uint32_t)c0, vz);
        tmem_ld_16x256b_x2_nowait(tl + kColIn + (uint32_t)c0, vi);
        tmem_ld_16x256b_x2_nowait(tl + kColHn + (uint32_t)c0, vh);
        tmem_ld_wait();
        reg_fence(vr); reg_fence(vz); reg_fence(vi); reg_fence(vh);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int c = c0 + 8 * i + 2 * fq;
          const float4 gb0 = S.c.gate_b[c], gb1 = S.c.gate_b[c + 1];
#pragma unroll
          for (int hr = 0; hr < 2; ++hr) {
            const int r = fr + 8 * hr;
            const uint32_t off = umma_off_bytes(r, c, H) >> 2;
            const float2 hh = *reinterpret_cast<const float2*>(S.hhi + off);
            const float2 hl = *reinterpret_cast<const float2*>(S.hlo + off);
            const int jj = 4 * i + 2 * hr;
            const float rg0 = sigmoid_fast(vr[jj] + gb0.x), rg1 = sigmoid_fast(vr[jj + 1] + gb1.x);
            const float zg0 = sigmoid_fast(vz[jj] + gb0.y), zg1 = sigmoid_fast(vz[jj + 1] + gb1.y);
            const float n0 = tanh_fast(vi[jj] + gb0.z + rg0 * (vh[jj] + gb0.w));
            const float n1 = tanh_fast(vi[jj + 1] + gb1.z + rg1 * (vh[jj + 1] + gb1.w));
            const float o0 = (1.0f - zg0) * n0 + zg0 * (hh.x + hl.x);
            const float o1 = (1.0f - zg1) * n1 + zg1 * (hh.y + hl.y);
            store_split2(S.hhi, S.hlo, r, c, H, o0, o1);
            if (r < valid) {
              const float2 v2 = make_float2(o0, o1);
              const size_t offg = (size_t)(row0 + r) * H + c;
              if (io.hidden_seq) *reinterpret_cast<float2*>(io.hidden_seq + tM * H + offg) = v2;
              if (io.hidden && t == T - 1) *reinterpret_cast<float2*>(io.hidden + offg) = v2;
            }
          }
        }
      }
      fence_async_smem();
      fence_before_sync();
      mbar_arrive(&S.a_ready);

      EP_STAMP(8);
      // ---- E5: Q tail, outputs, selection
      mbar_wait(&S.d_ready, d_par); d_par ^= 1u;
      fence_after_sync();
      EP_STAMP(9);
      const float bq2 = __ldg(W.bq2);
      {
        float acc[2][8], pa[2][8];
#pragma unroll
        for (int hr = 0; hr < 2; ++hr)
#pragma unroll
          for (int j = 0; j < 8; ++j) { acc[hr][j] = 0.f; pa[hr][j] = (j < A) ? Ps[j * kTcRows + fr + 8 * hr] : 0.f; }
        for (int cb = 0; cb < 4; ++cb) {
          float v[16];
          tmem_ld_16x256b_x4_nowait(tl + kColQ + (uint32_t)(32 * cb), v);
          tmem_ld_wait();
          reg_fence(v);
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int u = 32 * cb + 8 * i + 2 * fq + e;
              const float4 qc = S.c.q_c[u];
              const float4 wl = S.c.w1a[2 * u], wh = S.c.w1a[2 * u + 1];
              const float wa[8] = {wl.x, wl.y, wl.z, wl.w, wh.x, wh.y, wh.z, wh.w};
#pragma unroll
              for (int hr = 0; hr < 2; ++hr) {
                const float pre = v[4 * i + 2 * hr + e] + qc.x;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  acc[hr][j] = fmaf(qc.z, fmaxf(fmaf(pa[hr][j], qc.y, pre + wa[j]), 0.f), acc[hr][j]);
              }
            }
        }
#pragma unroll
        for (int hr = 0; hr < 2; ++hr)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float tot = acc[hr][j];
            tot += __shfl_xor_sync(0xffffffffu, tot, 1);
            tot += __shfl_xor_sync(0xffffffffu, tot, 2);
            if (fq == 0 && j < A) Qs[j * kTcRows + fr + 8 * hr] = tot + bq2;
          }
      }
      fence_before_sync();          // TMEM reads of this step are complete before the next x_full arrival
      __syncwarp();                 // Qs / Ps of a row were written by the lane with fq == 0 that owns it
      if (half == 0 && live) {
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
            const uint32_t row_id = (uint32_t)(io.rng_row_offset + row0 + r);
            const float u = io.u_eps ? io.u_eps[m] : philox_uniform(io.seed, kStreamEpsilon, row_id, (io.rng_step_dev ? __ldg(io.rng_step_dev) : io.rng_step) + t, 0);
            if (u < (io.epsilon_dev ? __ldg(io.epsilon_dev) : io.epsilon)) {
              if (io.rand_actions) {
                chosen = io.rand_actions[m];
              } else {
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
          const float pw = Ps[chosen * kTcRows + r];
          if (io.power) io.power[m] = pw;
          if (io.actions_mirror) io.actions_mirror[m] = chosen;      // e.g. the caller's page-locked host copy
          if (io.power_mirror) io.power_mirror[m] = pw;
          if (io.q_chosen) io.q_chosen[m] = Qs[chosen * kTcRows + r];
        }
      }
      __syncwarp();
      EP_STAMP(10);
    }
  }
  fence_before_sync();
  __syncthreads();
  TC_STAMP_ONCE(24);
  if (warp == 5) tmem_dealloc(tmem, kTmemCols);
  TC_STAMP_ONCE(25);
  TC_CTA_STAMP(1);
}

inline int agent_tc_chunk_k() { return kTcKc; }

inline size_t agent_tc_smem_bytes(const macjd_agent_weights& w) {
  return sizeof(TcSmem) + sizeof(float) * 2 * (size_t)w.n_actions * kTcRows + 1024;
}

inline bool agent_tc_supported(const macjd_agent_weights& w) {
  return w.tc_chunks != nullptr && w.hidden == kTcH && w.actor_hidden == kTcH && w.obs_pad % 32 == 0 &&
         w.n_actions <= 8 &&
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

inline int tc_profile_read(unsigned long long* out_host, int n) {
#ifdef MACJD_TC_PROFILE
  if (n > 64 + 1024) n = 64 + 1024;
  return cudaMemcpyFromSymbol(out_host, g_tc_prof, sizeof(unsigned long long) * n) == cudaSuccess ? MACJD_OK : MACJD_ERR_CUDA;
#else
  for (int i = 0; i < n; ++i) out_host[i] = 0;
  return MACJD_OK;
#endif
}

}  // namespace tc
}  // namespace macjd
#endif  // !MACJD_TEST_HOST_EMULATION
