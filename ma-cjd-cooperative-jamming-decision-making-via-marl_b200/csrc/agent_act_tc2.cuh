// Fused agent step on tcgen05 with CTA pairs (cta_group::2): 128 agent rows per 2-CTA cluster.
//
// Same computation as agent_forward_kernel (agent_act.cuh) with the dense layers on the tensor cores.
// Measurements on the B200 (tools/tc_mma_rate.py) show that one
// tcgen05.mma costs the same cycles for M = 64 and M = 128, so a CTA that owns 64 rows wastes
// half the pipe -- but 128 rows of TF32 hi + lo operand tiles do not fit one SM's shared
// memory.  A CTA pair solves both: each CTA keeps its own 64 rows of operands and only HALF of
// every weight chunk (the B operand of a 2-SM MMA is split across the pair), the leader CTA
// issues one M = 128, N = 128 tcgen05.mma.cta_group::2 per k-step and split product for both,
// and the weight bytes streamed from L2 per row halve.
//
// Per CTA (rank r of the pair): warps 0-7 epilogue, warp 8 weight stream (its half of each
// chunk), warp 9 = MMA issuer (leader: the whole warp runs the stage loop so that every operand is
// warp-uniform, one elected lane issues) or stage relay (peer: tells the leader when the peer's
// half has landed).  Accumulators: 2-SM M = 128 layout -- this CTA's row i on TMEM lane i for
// output units 0-63 and on lane 64 + i for units 64-127 (same columns) -- so epilogue warp w
// (TMEM lane quarter w % 4) owns rows 32 (w & 1) .. +31, the 64-unit half (w >> 1) & 1 and, of
// that half, units 32 (w >> 2) .. +31, with plain 32x32b loads; row sums (actor head, Q tail)
// are combined across the four threads of a row through shared memory.
// Cross-CTA events use cluster-scope mbarrier arrives (mapa) and multicast tcgen05.commit.
//
// Order inside a step (all seven accumulators live in separate TMEM columns):
//   issuer:   [actor.0 | fc1] x  ->  W_hr h, W_hz h, W_hn h  ->  actor.2 a1  ->  W_ir xf, W_iz xf, W_in xf  ->  q.0 h'
//   epilogue: x tile -> E1 (a1) .............. -> E3 (xf), E2 (actor head P) ....... -> E4 (gates, h') -> E5 (Q, selection)
#pragma once
#include "agent_tc_common.cuh"
#include "env_step2.cuh"
#ifndef MACJD_TEST_HOST_EMULATION
#include <type_traits>
#include <stdlib.h>
#include <cuda.h>            // CUtensorMap (type and enums only; the encoder is fetched through the runtime)
#endif

#ifndef MACJD_TEST_HOST_EMULATION
namespace macjd {
namespace tc {

constexpr int kT2Stages = 2;                                // ring stages in their own buffer ...
constexpr int kT2MaxStages = 4;                             // ... plus, where the xf tile is unused (io.part 4), its two halves
constexpr int kT2HalfBytes = kTcH * kTcKc * 4 / 2;        // this CTA's 64 weight rows of one hi (or lo) chunk
constexpr int kT2SubBytes = 2 * kT2HalfBytes;              // hi half + lo half of one chunk
constexpr int kT2StageBytes = 2 * kT2SubBytes;             // a stage holds two consecutive chunks: one barrier round trip per 64 k
constexpr int kT2ColSplit = 2;                             // epilogue warps per TMEM lane quarter (each takes 64 / split units)
constexpr int kT2EpiWarps = 4 * kT2ColSplit;
constexpr int kT2EpiThreads = 32 * kT2EpiWarps;
constexpr int kT2Threads = kT2EpiThreads + 64;             // + weight-stream warp + MMA / relay warp
constexpr int kT2EnvWarps = 4;                             // kFuseEnv launches: + the env-step warps (one group of lanes per env)
constexpr int kT2EnvThreads = 32 * kT2EnvWarps;
constexpr int kT2BarActions = 2, kT2BarEnvDone = 3;        // named barriers between the epilogue warps and the env warps
constexpr int kT2Upt = 64 / kT2ColSplit;                   // output units per epilogue thread
constexpr int kT2Parts = 2 * kT2ColSplit;                  // threads that share one row
// Accumulator columns (2-SM layout: 64 columns per 128-unit accumulator).  All seven layer
// accumulators of a step are live at once -- the recurrent products are issued before the
// actor's second layer and the actor head is evaluated while the input products run -- so none
// of them share columns, and Q has its own as well (448-511): in a launch that loops over the timesteps the issuer
// starts the NEXT step's observation and recurrent products while the epilogue still reads this step's Q.
constexpr uint32_t kT2ColA1 = 0, kT2ColFc1 = 64, kT2ColA2 = 128;
constexpr uint32_t kT2ColR = 192, kT2ColZ = 256, kT2ColIn = 320, kT2ColHn = 384, kT2ColQ = 448;
constexpr uint32_t kT2TmemCols = 512;
// Observation blocks ([64 rows][32 k], 8 KB per hi / lo part) are staged in up to four slots of the b0 tile (32 KB per part,
// free until E1 overwrites it with a1): the epilogue writes block after block and the issuer follows, instead of one
// write -> MMA -> "tile free" round trip per block (obs 176 = 6 blocks: 26 k of the step's 95 k cycles, tools/tc_phase_profile.py).
constexpr int kT2XSlots = 4;
constexpr int kT2XTileFloats = kTcRows * 32;
// Issue order of the eight K = 128 layers of a step, and where each sits in the packed weights
// (order there: 0 actor.2, 1 W_ir, 2 W_hr, 3 W_iz, 4 W_hz, 5 W_in, 6 W_hn, 7 q.0[:, :H]):
//   slot: 0 W_hr  1 W_hz  2 W_hn  3 actor.2  4 W_ir  5 W_iz  6 W_in  7 q.0
// The three recurrent products need nothing from this step's epilogues, so they run while the
// epilogue warps turn actor.0 into a1; the input products come after xf is written.
constexpr uint32_t kT2LayerSrc = 0x75310642u;    // nibble s = packed index of slot s
// io.part selects which slots a launch runs, in this order (nibble i = i-th slot):
//   0 whole step: 0..7;  1 recurrence only: the six GRU products;  2 heads only: q.0 (on the given
//   hidden state), then actor.2 -- one accumulator hand-over fewer, each gated by the epilogue.
//   3 input pre-pass: the three input products (each overwrites its accumulator);  4 recurrence on the
//   pre-computed input products: the three recurrent products, no observation stage.
__device__ __forceinline__ uint32_t t2_slot_seq(int part) {
  return part == 1 ? 0x00654210u : part == 2 ? 0x00000037u : part == 3 ? 0x00000654u : part == 4 ? 0x00000210u : 0x76543210u;
}
__device__ __forceinline__ int t2_slot_count(int part) { return part == 1 ? 6 : part == 2 ? 2 : (part == 3 || part == 4) ? 3 : 8; }

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 raddr;\n\t"
      "mapa.shared::cluster.u32 raddr, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [raddr];\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
// bounded wait with cluster-scope acquire (the barrier receives arrivals from the peer CTA)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
#pragma unroll 1   // (nvcc unrolls this spin loop 32x otherwise: a third of the pair kernel's code)
  for (uint32_t it = 0; it < (1u << 22); ++it) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
  }
  watchdog_raise();
}
__device__ __forceinline__ void mma_tf32_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of all prior MMAs of this thread -> the mbarrier at this offset in both CTAs
__device__ __forceinline__ void mma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .b16 mask;\n\t"
      "mov.b16 mask, 3;\n\t"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], mask;\n\t"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;\n" ::"n"(kT2EpiThreads) : "memory"); }

// Epilogue-side wait: one warp polls the mbarrier, the others park on a named barrier (a parked warp
// takes no issue slots).
__device__ __forceinline__ void epi_wait(uint64_t* bar, uint32_t parity, int warp) {
  if (warp == 0) mbar_wait_cluster(bar, parity);
  epi_bar_sync();
}

// Kernel parameters: the launch arguments plus a TMA descriptor of the packed weight chunks, seen as
// [bytes / 128][32 floats] so that one box of 64 rows is one contiguous 8 KB half-chunk.
struct T2Args {
  AgentArgs a;
  alignas(64) CUtensorMap wmap;
  Env2Args env;        // kFuseEnv launches: the environment step of the same timestep (macjd_rollout_step)
};

// 8 KB half-chunk -> this CTA's shared memory; the bytes are counted on the LEADER's mbarrier
// (cta_group::2 form: the mbarrier may live in the peer CTA), so the MMA issuer waits on one barrier
// for both halves of a stage and no thread has to relay "the peer's half has landed".
__device__ __forceinline__ void tma_half_chunk_2sm(void* smem_dst, const CUtensorMap* map, int row128, uint64_t* bar,
                                                   uint32_t leader_rank) {
  asm volatile(
      "{\n\t"
      ".reg .b32 rbar;\n\t"
      "mapa.shared::cluster.u32 rbar, %2, %5;\n\t"
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [rbar];\n\t"
      "}\n" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(0), "r"(row128), "r"(leader_rank)
      : "memory");
}

struct T2Smem {
  float b0hi[kTcRows * kTcH], b0lo[kTcRows * kTcH];        // observation block (first 8 KB) -> a1 -> xf
  float hhi[kTcRows * kTcH], hlo[kTcRows * kTcH];            // h -> h'
  unsigned char wst[kT2Stages][kT2StageBytes];
  TcConst c;
  float red[kT2Parts][8][kTcRows];          // partial row sums of the threads that share a row
  uint64_t w_full[kT2MaxStages], w_empty[kT2MaxStages];
  uint64_t x_full[kT2XSlots], x_empty[kT2XSlots];   // per observation slot: an mbarrier must not run two phases ahead of its waiter
  uint64_t d_ready, a_ready;
  uint64_t dx_ready;                        // acting launches: the observation products have completed (E1's wait)
  uint64_t t_full;                          // kBigA: the per-action tables have landed in the (dead) b0 tile
  uint64_t c_full;                          // the constant block has landed
  uint32_t tmem_base;
  int32_t act_s[kTcRows];                   // kFuseEnv: the actions just chosen, handed to the env step in place
  float pow_s[kTcRows];
};

// kWholeStep: the acting launch (io.part == 0) gets its own instance with the other parts compiled out -- less
// code to fetch on a cold start (the instruction fetches after an L2 flush are visible in the phase profile)
// kBigA: more than 8 discrete actions (BASELINE config 3: 2 x 16 radars + 1 = 33).  The per-action tables of the actor head
// and of the Q tail ([unit][A8] each, 20 KB at A = 33) then do not fit beside the operand tiles: they are read from global
// memory through L1 (warp-uniform 16-byte loads), the head / tail run in groups of 8 actions on register-resident
// activations, and the Q staging buffer borrows the xf operand tile, which is dead by then.
// kFuseEnv (acting launches only): the CTA also runs the environment step of its rows' envs -- a CTA's 64 rows are
// 64 / J whole envs.  The actions go from the selecting threads to the physics threads through shared memory (no
// second launch, no dependency hand-over, no global round trip); kT2EnvWarps extra warps, a group of lanes per env,
// evaluate the step on the derived scenario tables (env_step2.cuh: env2_physics_group, the code of the stand-alone
// kernel).  Nothing the agent reads at timestep t + 1 comes out of the physics of timestep t -- the views are the
// scenario's static rows (environment.py:479-551), copied by the epilogue threads while they would otherwise wait
// for the Q-head product -- so in a launch that loops over the timesteps the physics of step t (a dependent FP64
// chain of ~7 us) runs UNDER the agent's step t + 1 instead of between the two.
template <bool kWholeStep, bool kBigA, bool kFuseEnv = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kT2Threads + (kFuseEnv ? kT2EnvThreads : 0), 1)
agent_forward_tc2_kernel(const __grid_constant__ T2Args p) {
  extern __shared__ __align__(1024) unsigned char tc_raw[];
  T2Smem& S = *reinterpret_cast<T2Smem*>(tc_raw);
  const AgentArgs& a = p.a;
  const macjd_agent_weights& W = a.w;
  const macjd_agent_io& io = a.io;
  const int O = W.obs_dim, Op = W.obs_pad, A = W.n_actions;
  constexpr int H = kTcH;
  const int M = io.n_rows, T = io.n_steps;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  TC_CTA_STAMP(0);
  TC_STAMP_ONCE(20);
  const int row0 = blockIdx.x * kTcRows;
  const int valid = max(0, min(kTcRows, M - row0));
  const int A8 = (A + 7) & ~7;
  float* Ps = reinterpret_cast<float*>(tc_raw + sizeof(T2Smem));     // [A][64]
  // kBigA: the Q staging [A][64] borrows the xf tile (b0lo): every MMA that reads b0 has completed before E4, and
  // the next step's observation block is written only after E5
  // kBigA with A8 <= 48: the per-action tables ([H][A8] actor.4.weight^T, [H][A8] q.0 one-hot columns: 1 KB x A8) are copied
  // into the b0 tile once its last readers (the input products) have completed, and the actor head moves behind E4:
  // read through L1 they thrash it (the carve-out leaves ~30 KB of L1 beside 215 KB of shared memory; tools/tc_phase_profile.py
  // at obs 176 / 33 actions: head 22 k + tail 25 k cycles of a 95 k-cycle step, mostly L2 latency).  Qs sits behind them.
  const bool big_smem = kBigA && A8 <= 48;
  float* const tabs = S.b0hi;                                        // [H][A8] w3 then [H][A8] w1a (b0hi and b0lo are contiguous)
  float* Qs = kBigA ? (big_smem ? S.b0hi + 12288 : S.b0lo) : Ps + (size_t)A * kTcRows;
  const int nxc = Op / 32;
  const int chunks_per_step = 2 * kTcChunksPerX * nxc + 8 * kTcChunksPerH;
  const int mode = kWholeStep ? 0 : io.part;   // 0 whole step, 1 recurrence only, 2 heads only, 3 input pre-pass, 4 recurrence on gate_x
  const uint32_t slot_seq = t2_slot_seq(mode);
  const int nx = mode == 4 ? 0 : nxc;                            // observation stages per step
  const int supers_per_step = nx + 2 * t2_slot_count(mode);      // ring stages this launch runs per step
  // The recurrence on pre-computed input products (part 4) never touches the xf tile: its 64 KB become two more ring
  // stages.  A stage is refilled only after its MMAs have completed (commit -> TMA from L2, ~1.5 k cycles against
  // 0.9 k of MMA work per stage), so with two stages the tensor pipe waits for every refill; with four the stream
  // warp runs far enough ahead -- also across timesteps, the weights of step t + 1 being those of step t.
  const int nstages = (!kWholeStep && mode == 4) ? kT2MaxStages : kT2Stages;
  auto stage_base = [&](int s_) -> unsigned char* {
    return s_ < kT2Stages ? S.wst[s_] : (s_ == kT2Stages ? reinterpret_cast<unsigned char*>(S.b0hi) : reinterpret_cast<unsigned char*>(S.b0lo));
  };
  float* const xhi = S.b0hi;   // the observation block lives in b0 until E1 overwrites it with a1
  float* const xlo = S.b0lo;

  // Epilogue threads request their slice of the incoming hidden state and of the first observation
  // block right away, so those (cold) loads overlap the constant staging, the cluster barrier and
  // the TMEM allocation instead of following them.
  float4 h_in[kT2Upt / 4];
  float x_in[32 / kT2Parts];
  // Observation block: 64 rows x 8 float4 slots, two slots per epilogue thread.  When rows are 16-byte
  // multiples the slots are numbered in MEMORY order (a warp reads 512 contiguous bytes -- full lines
  // from HBM, full-size reads over PCIe when the observations sit in page-locked host memory); the
  // zero padding of k >= obs_dim takes the slots behind the data.  Otherwise thread = (row, 8 k) scalars.
  const size_t og = io.obs_group > 1 ? (size_t)io.obs_group : 1;   // rows [k og, (k + 1) og) share observation row k
  const bool x_vec = (O & 3) == 0;
  const int O4 = O >> 2;
  auto x_slot = [&](int s_, int xc_, int& r_, int& k4_) -> bool {
    const int w4 = min(8, O4 - 8 * xc_), nd = kTcRows * w4;
    if (s_ < nd) { r_ = s_ / w4; k4_ = s_ - r_ * w4; return true; }
    const int p_ = s_ - nd, wz = 8 - w4;
    r_ = p_ / wz; k4_ = w4 + (p_ - r_ * wz);
    return false;
  };
  // Observations of later steps of a kFuseEnv launch were written by this CTA's own view threads one step earlier:
  // they must not come through the non-coherent (read-only) path -- L2 loads (ld.global.cg) instead
  auto ld_obs4 = [](const float4* q_) -> float4 { return kFuseEnv ? __ldcg(q_) : __ldg(q_); };
  auto ld_obs1 = [](const float* q_) -> float { return kFuseEnv ? __ldcg(q_) : __ldg(q_); };
  float4 x4_in[2];
  if (warp < kT2EpiWarps && x_vec) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int r_, k4_;
      const bool data = x_slot(tid + i * kT2EpiThreads, 0, r_, k4_);
      x4_in[i] = (data && r_ < valid && mode != 4)
                     ? __ldg(reinterpret_cast<const float4*>(io.obs + ((size_t)(row0 + r_) / og) * O) + k4_) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  if (warp < kT2EpiWarps) {
    const int q4 = warp & 3, ch = warp >> 2, half = q4 >> 1;
    const int r = (q4 & 1) * 32 + lane, ub = half * 64 + ch * kT2Upt, part = half * kT2ColSplit + ch;
    const bool live = r < valid;
    const float* hsrc = io.hidden_in ? io.hidden_in : io.hidden;      // initial state: a separate read-only source, or in place
    const bool have = hsrc && !io.hidden_zero_init && live;
#pragma unroll
    for (int i = 0; i < kT2Upt / 4; ++i)
      h_in[i] = have ? __ldg(reinterpret_cast<const float4*>(hsrc + (size_t)(row0 + r) * H + ub) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 32 / kT2Parts; ++j) {
      const int kk = part * (32 / kT2Parts) + j;
      x_in[j] = (live && kk < O && mode != 4 && !x_vec) ? __ldg(io.obs + ((size_t)(row0 + r) / og) * O + kk) : 0.f;
    }
  }

  const float bq2 = __ldg(W.bq2);   // used at the very end (E5): requested here so that a cold line is not a stall there
  // kBigA: per-action tables behind the constant block: [H][A8] actor.4.weight^T, [H][A8] q.0 one-hot columns, [A8] actor.4.bias
  const float* const big_w3 = W.tc_chunks + (size_t)chunks_per_step * kTcChunkFloats + sizeof(TcConst) / 4;
  const float* const big_w1a = big_w3 + (size_t)H * A8;
  const float* const big_b3 = big_w1a + (size_t)H * A8;
  TC_STAMP_ONCE(26);
  if (tid < kT2Threads) warm_weights_l2(W.tc_chunks, chunks_per_step, kT2Threads);
  TC_STAMP_ONCE(27);

  if (tid == 0) {
    // "full" lives in the leader: armed by its stream thread for both CTAs' bytes of a stage
    for (int s = 0; s < kT2MaxStages; ++s) { mbar_init(&S.w_full[s], 1); mbar_init(&S.w_empty[s], 1); }
    for (int s = 0; s < kT2XSlots; ++s) { mbar_init(&S.x_full[s], 2 * kT2EpiThreads); mbar_init(&S.x_empty[s], 1); }
    mbar_init(&S.d_ready, 1); mbar_init(&S.a_ready, 2 * kT2EpiThreads);
    mbar_init(&S.dx_ready, 1);
    mbar_init(&S.t_full, 1);
    mbar_init(&S.c_full, 1);
    fence_mbar_init();
    // the per-layer vectors (TcConst, packed by the host behind the weight chunks): one 13.9 KB bulk copy
    // instead of ~26 scalar loads and stores per thread (1.25 k cycles warm, 3 k after an L2 flush)
    mbar_expect_tx(&S.c_full, (uint32_t)sizeof(TcConst));
    bulk_g2s(&S.c, reinterpret_cast<const char*>(W.tc_chunks) + (size_t)chunks_per_step * kTcChunkBytes, (uint32_t)sizeof(TcConst),
             &S.c_full);
  }
  // tensor memory for the pair (needs nothing from the rest of the prologue: runs under the constant staging)
  TC_STAMP_ONCE(28);
  if (warp == kT2EpiWarps + 1) tmem_alloc_2sm(&S.tmem_base, kT2TmemCols);
  TC_STAMP_ONCE(21);
  // one cluster-wide barrier ends the prologue: both CTAs' mbarriers exist before any remote arrive, the
  // staged constants and the TMEM base address are visible to every warp
  fence_before_sync();
  cluster_sync_all();
  fence_after_sync();
  TC_STAMP_ONCE(22);
  mbar_wait(&S.c_full, 0);                     // constants in place (every thread that reads them waits here)
  const uint32_t tmem = S.tmem_base;
  TC_STAMP_ONCE(23);
  // the next kernel of the stream may be scheduled now if it asked for an early start (the env step does:
  // its launch latency and table loads then overlap this kernel; it waits before reading the actions)
  grid_launch_dependents();

  if (warp == kT2EpiWarps) {
    // =========================================================== weight stream: this CTA's half of every chunk
    if (lane == 0) {
      uint32_t empty_par = 0;
      int s = 0;
      for (int t = 0; t < T; ++t) {
        for (int L = 0; L < supers_per_step; ++L) {
          if (t > 0 || L >= nstages) { mbar_wait_cluster(&S.w_empty[s], (empty_par >> s) & 1u); empty_par ^= 1u << s; }
          // stages follow the issue order; the packed buffer keeps the single-CTA kernel's layer order
          const int slot = L < nx ? 0 : (int)((slot_seq >> (4 * ((L - nx) >> 1))) & 0xFu);
          const int Lsrc = L < nx ? L : nxc + 2 * (int)((kT2LayerSrc >> (4 * slot)) & 0xFu) + ((L - nx) & 1);
          // 128-byte row index of this CTA's half of the stage's first chunk
          const int row = (int)(((size_t)Lsrc * 2 * kTcChunkBytes + (size_t)rank * kT2HalfBytes) / 128);
          if (rank == 0) mbar_expect_tx(&S.w_full[s], 2 * kT2StageBytes);       // both CTAs' halves
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            tma_half_chunk_2sm(stage_base(s) + sub * kT2SubBytes, &p.wmap, row + sub * (kTcChunkBytes / 128), &S.w_full[s], 0);                   // hi
            tma_half_chunk_2sm(stage_base(s) + sub * kT2SubBytes + kT2HalfBytes, &p.wmap, row + sub * (kTcChunkBytes / 128) + kTcChunkBytes / 256,
                               &S.w_full[s], 0);                                                                                                      // lo
          }
          s = (s + 1 == nstages) ? 0 : s + 1;
        }
      }
    }
  } else if (warp == kT2EpiWarps + 1) {
    if (rank != 0) {
      // the peer's MMA warp has nothing to do: its loads report to the leader's barriers by themselves
    } else if (rank == 0) {
      // =========================================================== leader: MMA issue for the pair
      // The WHOLE warp runs this loop and one elected lane issues: every MMA operand is then a
      // warp-uniform value held in uniform registers.  Issued from inside `if (lane == 0)` each
      // tcgen05.mma was wrapped in an ELECT / 5x R2UR / branch sequence (60 cycles per MMA).
      // The schedule is a compact loop over stages (a stage = two consecutive 32-k weight chunks
      // against one activation tile) with the per-stage parameters computed from the stage index:
      // straight-line code for the 408 MMAs of a step is ~100 KB of instructions executed once per
      // step, and after an L2 flush the issuer then runs at instruction-fetch speed (+23 us).
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
      const uint32_t idesc = umma_idesc_tf32(2 * kTcRows, H);
      const uint32_t bh = smem_u32(S.b0hi), bl = smem_u32(S.b0lo), hh = smem_u32(S.hhi), hl = smem_u32(S.hlo);
      const uint32_t wb[kT2MaxStages] = {smem_u32(stage_base(0)), smem_u32(stage_base(1)), smem_u32(stage_base(2)), smem_u32(stage_base(3))};
      uint32_t full_par = 0, x_full_par = 0, a_ready_par = 0;
      int s = 0;
      for (int t = 0; t < T; ++t) {
#ifdef MACJD_TC_PROFILE
        int stamp = 33;
#endif
        for (int L = 0; L < supers_per_step; ++L) {
          uint32_t ahi, alo, sbo, koff0, koff1, d0, d1, first0, first1, pre, post;
          if (L < nx) {                        // [actor.0 | fc1] chunk pair of observation block L
            ahi = bh + (uint32_t)(L & (kT2XSlots - 1)) * (kT2XTileFloats * 4); alo = bl + (uint32_t)(L & (kT2XSlots - 1)) * (kT2XTileFloats * 4);
            sbo = 32 * 32; koff0 = koff1 = 0;
            d0 = kT2ColA1; d1 = kT2ColFc1; first0 = first1 = (L == 0);
            pre = 1; post = (L + kT2XSlots < nx ? 1u : 0u) | (L == nx - 1 ? (kWholeStep ? 4u : 2u) : 0u);   // (x_empty: the slot is reused)
          } else {                             // K = 128 layers: 4 chunks = 2 stages each
            //  slot j:  0 W_hr (h)  1 W_hz (h)  2 W_hn (h)  3 actor.2 (b0)  4 W_ir (b0)  5 W_iz (b0)  6 W_in (b0)  7 q.0 (h)
            const uint32_t j = (slot_seq >> (4 * ((uint32_t)(L - nx) >> 1))) & 0xFu, hf = (uint32_t)(L - nx) & 1u;
            const bool use_h = (0x87u >> j) & 1u;          // slots 0, 1, 2, 7 read h
            ahi = use_h ? hh : bh; alo = use_h ? hl : bl; sbo = H * 32;
            koff0 = (2 * hf) * kTcAStep; koff1 = (2 * hf + 1) * kTcAStep;
            // accumulator column / 64 per slot: R 3, Z 4, Hn 6, A2 2, R 3, Z 4, In 5, Q 7
            // (heads only: q.0 is issued before actor.0's accumulator has been read, so it takes W_hn's columns)
            d0 = d1 = 64u * (((mode == 2 ? 0x65432643u : 0x75432643u) >> (4 * j)) & 0xFu);
            // W_ir, W_iz accumulate onto the recurrent product (pre-pass: they stand alone)
            first0 = (((mode == 3 ? 0xFFu : 0xCFu) >> j) & 1u) & (hf == 0 ? 1u : 0u);
            first1 = 0;
            // whole step: actor.2 (a1), W_ir (xf), q.0 (h') wait for the epilogue's tile; after actor.2, W_in,
            // q.0 the epilogue may read.  Heads only: q.0 reads the given hidden state (no wait, no hand-over
            // of its own: actor.2's commit covers it).
            // Recurrence on gate_x: W_hr waits for h' of the previous step, W_hn's commit hands over.
            const uint32_t pre_m = mode == 2 ? 0x08u : mode == 4 ? 0x01u : 0x98u;
            const uint32_t post_m = mode == 2 ? 0x08u : mode == 4 ? 0x04u : 0xC8u;
            pre = (hf == 0 && ((pre_m >> j) & 1u)) ? 2u : 0u;
            post = (hf == 1 && ((post_m >> j) & 1u)) ? 2u : 0u;
          }
          const int xs = L & (kT2XSlots - 1);       // observation slot of this stage (pre == 1 only)
          if (pre == 1) { mbar_wait_cluster(&S.x_full[xs], (x_full_par >> xs) & 1u); x_full_par ^= 1u << xs; if (lane == 0) TC_STAMP(32); }
          if (pre == 2) {
            mbar_wait_cluster(&S.a_ready, a_ready_par); a_ready_par ^= 1u;
#ifdef MACJD_TC_PROFILE
            if (lane == 0) TC_STAMP(stamp);
            ++stamp;
#endif
          }
          mbar_wait_cluster(&S.w_full[s], (full_par >> s) & 1u);
          full_par ^= 1u << s;
          fence_after_sync();
          const uint32_t wbase = s == 0 ? wb[0] : s == 1 ? wb[1] : s == 2 ? wb[2] : wb[3];
          if (elect_one()) {
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
              const uint32_t koff = sub ? koff1 : koff0;
              const uint64_t dah = umma_smem_desc(ahi + koff, 128, sbo);
              const uint64_t dal = umma_smem_desc(alo + koff, 128, sbo);
              const uint64_t dbh = umma_smem_desc(wbase + sub * kT2SubBytes, 128, kTcKc * 32);
              const uint64_t dbl = umma_smem_desc(wbase + sub * kT2SubBytes + kT2HalfBytes, 128, kTcKc * 32);
              const uint32_t d = tmem_u + (sub ? d1 : d0);
              const uint32_t first = sub ? first1 : first0;
#pragma unroll
              for (int ks = 0; ks < kTcKc / 8; ++ks) {
                const uint64_t adv = (uint64_t)((ks * 256) >> 4);
                mma_tf32_ss_2sm(d, dah + adv, dbh + adv, idesc, (first && ks == 0) ? 0u : 1u);
                mma_tf32_ss_2sm(d, dal + adv, dbh + adv, idesc, 1u);
                mma_tf32_ss_2sm(d, dah + adv, dbl + adv, idesc, 1u);
              }
            }
            mma_commit_2sm(&S.w_empty[s]);
            if (post & 1u) mma_commit_2sm(&S.x_empty[xs]);
            if (post & 2u) mma_commit_2sm(&S.d_ready);
            if (post & 4u) mma_commit_2sm(&S.dx_ready);
          }
          __syncwarp();
#ifdef MACJD_TC_PROFILE
          if (post & 6u) { if (lane == 0) TC_STAMP(stamp); ++stamp; }
#endif
          s = (s + 1 == nstages) ? 0 : s + 1;
        }
      }
    }
  } else if (kFuseEnv && warp >= kT2EpiWarps + 2) {
    // =========================================================== env warps: the step of this CTA's envs, one timestep
    // behind the agent (environment.py:221-477).  Group `slot` of G lanes owns env row0 / J + slot for the whole launch.
    const Env2Args& E = p.env;
    const int Je = E.tab.n_jammers, epc = kTcRows / Je, G = E.group;
    const int passes = (epc * G + kT2EnvThreads - 1) / kT2EnvThreads;   // (whole warps: epc * G is a multiple of 32)
    // pull the envs' derived blocks into L1 / L2 while the agent's first step runs (read-only for the whole launch)
    if (E.tab.env_stride != 0) {
      const double* blk = E.tab.derived + (int64_t)(row0 / Je) * E.rows.total;
      double warm = 0.0;
      for (int off = (tid - kT2Threads) * 4; off < (valid / Je) * E.rows.total; off += kT2EnvThreads * 4) warm += __ldg(blk + off);
      asm volatile("" ::"d"(warm));
    }
    for (int t = 0; t < T; ++t) {
      asm volatile("bar.sync %0, %1;\n" ::"n"(kT2BarActions), "n"(kT2EpiThreads + kT2EnvThreads) : "memory");   // step t's actions are staged
      const macjd_env_io eio = env2_io_at(E, t);
      if (tid == kT2Threads) TC_STAMP(40);
      for (int ps = 0; ps < passes; ++ps) {
        const int et = tid - kT2Threads + ps * kT2EnvThreads;
        if (et - (et & 31) >= epc * G) break;         // warp-uniform
        const int slot = et / G, g = et - slot * G;
        const bool env_live = slot < valid / Je;
        const int e = row0 / Je + (env_live ? slot : 0);
        const double* d = E.tab.derived + (E.tab.env_stride == 0 ? (int64_t)0 : (int64_t)e * E.rows.total);
        env2_physics_group(E, eio, e, env_live, g, G, d, S.act_s + slot * Je, S.pow_s + slot * Je);
      }
      if (tid == kT2Threads) TC_STAMP(41);
      if (t + 1 < T) {                                // the staging buffers may take the next step's actions
        __threadfence_block();
        asm volatile("bar.arrive %0, %1;\n" ::"n"(kT2BarEnvDone), "n"(kT2EpiThreads + kT2EnvThreads) : "memory");
      }
    }
  } else {
    // =========================================================== epilogue warps
    // warp -> TMEM lane quarter q = warp % 4 (a hardware rule) = (row block, 64-unit half of the 2-SM
    // accumulator layout); warps q and q + 4 split that quarter's 64 columns
    const int q4 = warp & 3, ch = warp >> 2;
    const int half = q4 >> 1;
    const int r = (q4 & 1) * 32 + lane;              // row within this CTA's tile
    const int ub = half * 64 + ch * kT2Upt;          // first output unit of this thread
    const int part = half * kT2ColSplit + ch;        // which of the row's kT2Parts threads
    const bool live = r < valid;
    const uint32_t tl = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(ch * kT2Upt);
    uint32_t d_par = 0, x_empty_par = 0, dx_par = 0, t_par = 0;
    // kBigA: one bulk copy brings both per-action tables into the b0 tile (call once its MMA readers have completed)
    auto stage_tables = [&]() {
      if (big_smem && tid == 0) {
        fence_async_smem();
        mbar_expect_tx(&S.t_full, (uint32_t)(1024 * A8));
        bulk_g2s(tabs, big_w3, (uint32_t)(1024 * A8), &S.t_full);
      }
    };
    // kFuseEnv launches whose observation is one 32-wide block: the next step's observation tile is staged during THIS
    // step's E5 (below), so the issuer runs the next step's observation and recurrent products under E5 / E1
    const bool early_x = kFuseEnv && !kBigA && nx == 1 && x_vec;

#pragma unroll
    for (int i = 0; i < kT2Upt / 4; ++i) {
      const float v[4] = {h_in[i].x, h_in[i].y, h_in[i].z, h_in[i].w};
      store_split4(S.hhi, S.hlo, r, ub + 4 * i, H, v);
    }
    if (mode == 4) {                    // no observation stage announces the hidden tile: say so directly
      fence_async_smem();
      fence_before_sync();
      mbar_arrive_cluster(&S.a_ready, 0);
    }

    for (int t = 0; t < T; ++t) {
      const size_t tM = (size_t)t * M;
      EP_STAMP(0);
      // this step's availability mask, requested now and used in E5 (a cold load at the end of the
      // step would sit on the critical path)
      uint64_t av_mask = ~0ull;
      if (part == 0 && live && io.avail) {
        const uint8_t* ap = io.avail + (tM + row0 + r) * A;
        av_mask = 0;
        if (kBigA) {
          // A <= 64 bytes per row at an arbitrary alignment: the aligned 4-byte words that cover the row, all requested at
          // once (one byte load per action was a chain of ~8 dependent memory round trips at the head of the step);
          // the last row of the buffer is read bytewise so that no word reaches past the caller's allocation
          const bool last_row = (tM + row0 + r + 1) == (size_t)T * M;
          if (!last_row) {
            const uintptr_t a0 = reinterpret_cast<uintptr_t>(ap);
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(a0 & ~(uintptr_t)3);
            const int sh = (int)(a0 & 3), nw = (sh + A + 3) >> 2;
            uint32_t wd[17];
#pragma unroll
            for (int i = 0; i < 17; ++i) wd[i] = i < nw ? (kFuseEnv ? __ldcg(wp + i) : __ldg(wp + i)) : 0u;
#pragma unroll
            for (int i = 0; i < 17; ++i)
#pragma unroll
              for (int b_ = 0; b_ < 4; ++b_) {
                const int act = 4 * i + b_ - sh;
                if (act >= 0 && act < A && ((wd[i] >> (8 * b_)) & 0xFFu) != 0) av_mask |= 1ull << act;
              }
          } else {
#pragma unroll 4
            for (int act = 0; act < A; ++act) av_mask |= (uint64_t)((kFuseEnv ? __ldcg(ap + act) : __ldg(ap + act)) != 0 ? 1u : 0u) << act;
          }
        } else {
#pragma unroll
          for (int act = 0; act < 8; ++act)
            if (act < A) av_mask |= (uint64_t)((kFuseEnv ? __ldcg(ap + act) : __ldg(ap + act)) != 0 ? 1u : 0u) << act;
        }
      }
      // the first slots' worth of blocks: every load is requested before the first store
      constexpr int kPre = kFuseEnv ? 1 : kT2XSlots;      // (the fused instance runs at a 128-register cap)
      float4 x4_blk[kPre][2];
      if (x_vec && !(early_x && t > 0)) {
#pragma unroll
        for (int xc = 0; xc < kPre; ++xc)
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            int r_, k4_;
            const bool data = xc < nx && x_slot(tid + i * kT2EpiThreads, xc, r_, k4_);
            x4_blk[xc][i] = (t == 0 && xc == 0) ? x4_in[i]          // (requested at kernel entry)
                            : (data && r_ < valid) ? ld_obs4(reinterpret_cast<const float4*>(io.obs + ((tM + row0 + r_) / og) * O + 32 * xc) + k4_)
                                                   : make_float4(0.f, 0.f, 0.f, 0.f);
          }
      }
      for (int xc = 0; xc < nx; ++xc) {
        if (early_x && t > 0) break;             // staged one step ahead
        // a slot is reused by block xc + kT2XSlots of the same step: wait for block xc's products (the issuer commits
        // x_empty for exactly those blocks).  Across steps nothing is needed: every later wait of the epilogue is on a
        // commit issued behind the observation products.
        const int xs = xc & (kT2XSlots - 1);
        if (xc >= kT2XSlots) { epi_wait(&S.x_empty[xs], (x_empty_par >> xs) & 1u, warp); x_empty_par ^= 1u << xs; }
        float* const sxhi = xhi + (xc & (kT2XSlots - 1)) * kT2XTileFloats;
        float* const sxlo = xlo + (xc & (kT2XSlots - 1)) * kT2XTileFloats;
        const float* obs = io.obs + ((tM + row0 + r) / og) * O;
        const bool pre = (t == 0 && xc == 0);      // already in registers (requested at kernel entry)
        if (x_vec) {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            int r_, k4_;
            const bool data = x_slot(tid + i * kT2EpiThreads, xc, r_, k4_);
            float4 q;
            if (xc < kPre) {
              q = x4_blk[0][i];
#pragma unroll
              for (int b_ = 1; b_ < kPre; ++b_) q = xc == b_ ? x4_blk[b_][i] : q;
            } else {
              q = (data && r_ < valid) ? ld_obs4(reinterpret_cast<const float4*>(io.obs + ((tM + row0 + r_) / og) * O + 32 * xc) + k4_)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            const float v[4] = {q.x, q.y, q.z, q.w};
            store_split4(sxhi, sxlo, r_, 4 * k4_, 32, v);
          }
        } else
#pragma unroll
        for (int g = 0; g < 32 / kT2Parts / 4; ++g) {
          const int k = part * (32 / kT2Parts) + 4 * g;
          float v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int kk = xc * 32 + k + j;
            v[j] = pre ? x_in[4 * g + j] : ((live && kk < O) ? ld_obs1(obs + kk) : 0.f);
          }
          store_split4(sxhi, sxlo, r, k, 32, v);
        }
        fence_async_smem();
        fence_before_sync();
        mbar_arrive_cluster(&S.x_full[xs], 0);
      }

      if (mode == 0 || mode == 2) {   // (recurrence / pre-pass launches have no actor)
      // ---- E1: a1 = relu(D1 + b) -> B0
      EP_STAMP(1);
      if (kWholeStep) { epi_wait(&S.dx_ready, dx_par, warp); dx_par ^= 1u; }
      else { epi_wait(&S.d_ready, d_par, warp); d_par ^= 1u; }
      fence_after_sync();
      EP_STAMP(2);
      for (int c0 = 0; c0 < kT2Upt; c0 += 16) {
        float v[16];
        tmem_ld16_nowait(tl + kT2ColA1 + (uint32_t)c0, v);
        tmem_ld_wait();
        reg_fence(v);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = fmaxf(v[4 * q + j] + S.c.ba1[ub + c0 + 4 * q + j], 0.f);
          store_split4(S.b0hi, S.b0lo, r, ub + c0 + 4 * q, H, o);
        }
      }
      fence_async_smem();
      fence_before_sync();
      mbar_arrive_cluster(&S.a_ready, 0);
      EP_STAMP(3);
      }

      // ---- E3: xf = relu(D3 + b) -> B0 (releases the input products), then E2: actor head while they run.
      // This wait is for actor.2 (whole step, heads only) or for the observation products (recurrence only).
      if (mode != 4) { epi_wait(&S.d_ready, d_par, warp); d_par ^= 1u; }
      fence_after_sync();
      EP_STAMP(4);
      if (kBigA && mode == 2) stage_tables();          // heads only: a1's last reader (actor.2) has completed
      if (mode != 2 && mode != 4) {
      for (int c0 = 0; c0 < kT2Upt; c0 += 16) {
        float v[16];
        tmem_ld16_nowait(tl + kT2ColFc1 + (uint32_t)c0, v);
        tmem_ld_wait();
        reg_fence(v);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = fmaxf(v[4 * q + j] + S.c.bfc1[ub + c0 + 4 * q + j], 0.f);
          store_split4(S.b0hi, S.b0lo, r, ub + c0 + 4 * q, H, o);
        }
      }
      fence_async_smem();
      fence_before_sync();
      mbar_arrive_cluster(&S.a_ready, 0);
      }
      EP_STAMP(5);
      if ((mode == 0 || mode == 2) && !kBigA)
      {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int c0 = 0; c0 < kT2Upt; c0 += 16) {
          float v[16];
          tmem_ld16_nowait(tl + kT2ColA2 + (uint32_t)c0, v);
          tmem_ld_wait();
          reg_fence(v);
#pragma unroll
          for (int n = 0; n < 16; ++n) {
            const int u = ub + c0 + n;
            const float a2 = fmaxf(v[n] + S.c.ba2[u], 0.f);
            const float4 wl = S.c.wa3t[2 * u], wh = S.c.wa3t[2 * u + 1];
            acc[0] = fmaf(a2, wl.x, acc[0]); acc[1] = fmaf(a2, wl.y, acc[1]);
            acc[2] = fmaf(a2, wl.z, acc[2]); acc[3] = fmaf(a2, wl.w, acc[3]);
            acc[4] = fmaf(a2, wh.x, acc[4]); acc[5] = fmaf(a2, wh.y, acc[5]);
            acc[6] = fmaf(a2, wh.z, acc[6]); acc[7] = fmaf(a2, wh.w, acc[7]);
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) S.red[part][j][r] = acc[j];
        epi_bar_sync();
        if (part == 0)
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (j < A) {
              float sum = S.c.ba3[j];
#pragma unroll
              for (int pp = 0; pp < kT2Parts; ++pp) sum += S.red[pp][j][r];
              Ps[j * kTcRows + r] = sigmoid_fast(sum);
            }
      }
      // kBigA actor head: P_a = sigmoid(a2 . actor.4.weight[a] + b) for A8 actions, from the A2 accumulator
      auto actor_head_big = [&](const float4* w3, auto ld) {
        // this thread's 32 units of a2 stay in registers; 8 actions per pass over them
        float a2v[kT2Upt];
#pragma unroll
        for (int c0 = 0; c0 < kT2Upt; c0 += 16) {
          float v[16];
          tmem_ld16_nowait(tl + kT2ColA2 + (uint32_t)c0, v);
          tmem_ld_wait();
          reg_fence(v);
#pragma unroll
          for (int n = 0; n < 16; ++n) a2v[c0 + n] = fmaxf(v[n] + S.c.ba2[ub + c0 + n], 0.f);
        }
        const int a84 = A8 >> 2;
#pragma unroll 1
        for (int g = 0; g < A8; g += 8) {
          float acc[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
          for (int n = 0; n < kT2Upt; ++n) {
            const float4 wl = ld(w3 + n * a84 + (g >> 2)), wh = ld(w3 + n * a84 + (g >> 2) + 1);
            const float a2 = a2v[n];
            acc[0] = fmaf(a2, wl.x, acc[0]); acc[1] = fmaf(a2, wl.y, acc[1]);
            acc[2] = fmaf(a2, wl.z, acc[2]); acc[3] = fmaf(a2, wl.w, acc[3]);
            acc[4] = fmaf(a2, wh.x, acc[4]); acc[5] = fmaf(a2, wh.y, acc[5]);
            acc[6] = fmaf(a2, wh.z, acc[6]); acc[7] = fmaf(a2, wh.w, acc[7]);
          }
          if (g > 0) epi_bar_sync();         // the previous group's partial sums have been consumed
#pragma unroll
          for (int j = 0; j < 8; ++j) S.red[part][j][r] = acc[j];
          epi_bar_sync();
          // the row's four threads finish two actions each
#pragma unroll
          for (int jj = 0; jj < 8 / kT2Parts; ++jj) {
            const int j = part * (8 / kT2Parts) + jj;
            if (g + j < A) {
              float sum = __ldg(big_b3 + g + j);
#pragma unroll
              for (int pp = 0; pp < kT2Parts; ++pp) sum += S.red[pp][j][r];
              Ps[(g + j) * kTcRows + r] = sigmoid_fast(sum);
            }
          }
        }
      };
      if ((mode == 0 || mode == 2) && kBigA && !big_smem)
        actor_head_big(reinterpret_cast<const float4*>(big_w3 + (size_t)ub * A8), [](const float4* q_) { return __ldg(q_); });
      EP_STAMP(6);
      float4 x4_next[2];
      if (kFuseEnv) {
        // While the input products run (this wait used to be idle): the next timestep's state / obs / avail of this
        // CTA's envs -- nothing of them depends on the actions (environment.py:479-551: static views) -- and the loads
        // of the next step's observation tile.  What the agent observes at t + 1 is what env2_views writes: the env's
        // static state row, once per jammer (environment.py:512-522); the tile is filled from the same source rows.
        const int Je = p.env.tab.n_jammers, ne = valid / Je;
        if (early_x && t + 1 < T) {
          const bool shared_scn = p.env.tab.env_stride == 0;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            int r_, k4_;
            const bool data = x_slot(tid + i * kT2EpiThreads, 0, r_, k4_);
            x4_next[i] = (data && r_ < valid)
                             ? __ldg(reinterpret_cast<const float4*>(p.env.state_rows + (shared_scn ? (size_t)0 : (size_t)((row0 + r_) / Je) * O)) + k4_)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        env2_views(p.env, env2_io_at(p.env, t), row0 / Je, ne, tid, kT2EpiThreads);
      }

      if (mode == 3) {
      // ---- pre-pass: the three input products of every row -> gate_x [row][3][H]
      epi_wait(&S.d_ready, d_par, warp); d_par ^= 1u;
      fence_after_sync();
      for (int c0 = 0; c0 < kT2Upt; c0 += 8) {
        float vr[8], vz[8], vi[8];
        tmem_ld8_nowait(tl + kT2ColR + (uint32_t)c0, vr);
        tmem_ld8_nowait(tl + kT2ColZ + (uint32_t)c0, vz);
        tmem_ld8_nowait(tl + kT2ColIn + (uint32_t)c0, vi);
        tmem_ld_wait();
        reg_fence(vr); reg_fence(vz); reg_fence(vi);
        if (live) {
          float4* g = reinterpret_cast<float4*>(io.gate_x + (size_t)(row0 + r) * 3 * H + ub + c0);
          g[0] = make_float4(vr[0], vr[1], vr[2], vr[3]); g[1] = make_float4(vr[4], vr[5], vr[6], vr[7]);
          g[H / 4] = make_float4(vz[0], vz[1], vz[2], vz[3]); g[H / 4 + 1] = make_float4(vz[4], vz[5], vz[6], vz[7]);
          g[2 * H / 4] = make_float4(vi[0], vi[1], vi[2], vi[3]); g[2 * H / 4 + 1] = make_float4(vi[4], vi[5], vi[6], vi[7]);
        }
      }
      fence_before_sync();
      } else if (mode != 2) {
      // ---- E4: GRU gates -> h' (in place over h), global hidden outputs
      // mode 4: the input products come from gate_x (software-pipelined: the next 8 units' values are
      // requested while the current ones are used; the first request goes out before the wait)
      const float4* gx = reinterpret_cast<const float4*>(io.gate_x + (tM + row0 + r) * 3 * H + ub);
      float4 gq[6];
      if (mode == 4) {
#pragma unroll
        for (int i = 0; i < 6; ++i) gq[i] = live ? __ldg(gx + (i >> 1) * (H / 4) + (i & 1)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      epi_wait(&S.d_ready, d_par, warp); d_par ^= 1u;
      fence_after_sync();
      EP_STAMP(7);
      if (kBigA && mode == 0) stage_tables();          // xf's last readers (the input products) have completed
      for (int c0 = 0; c0 < kT2Upt; c0 += 8) {
        float vr[8], vz[8], vi[8], vh[8];
        tmem_ld8_nowait(tl + kT2ColR + (uint32_t)c0, vr);
        tmem_ld8_nowait(tl + kT2ColZ + (uint32_t)c0, vz);
        if (mode != 4) tmem_ld8_nowait(tl + kT2ColIn + (uint32_t)c0, vi);
        tmem_ld8_nowait(tl + kT2ColHn + (uint32_t)c0, vh);
        tmem_ld_wait();
        reg_fence(vr); reg_fence(vz); reg_fence(vh);
        if (mode != 4) {
          reg_fence(vi);
        } else {
          const float gr_[8] = {gq[0].x, gq[0].y, gq[0].z, gq[0].w, gq[1].x, gq[1].y, gq[1].z, gq[1].w};
          const float gz_[8] = {gq[2].x, gq[2].y, gq[2].z, gq[2].w, gq[3].x, gq[3].y, gq[3].z, gq[3].w};
          const float gn_[8] = {gq[4].x, gq[4].y, gq[4].z, gq[4].w, gq[5].x, gq[5].y, gq[5].z, gq[5].w};
#pragma unroll
          for (int i = 0; i < 8; ++i) { vr[i] += gr_[i]; vz[i] += gz_[i]; vi[i] = gn_[i]; }
          if (c0 + 8 < kT2Upt) {
#pragma unroll
            for (int i = 0; i < 6; ++i)
              gq[i] = live ? __ldg(gx + (i >> 1) * (H / 4) + (c0 + 8) / 4 + (i & 1)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int c = ub + c0 + 4 * q;
          const uint32_t off = umma_off_bytes(r, c, H) >> 2;
          const float4 hh = *reinterpret_cast<const float4*>(S.hhi + off);
          const float4 hl = *reinterpret_cast<const float4*>(S.hlo + off);
          const float hold[4] = {hh.x + hl.x, hh.y + hl.y, hh.z + hl.z, hh.w + hl.w};
          float o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int u = c + j, jj = 4 * q + j;
            const float4 gb = S.c.gate_b[u];
            const float rg = sigmoid_fast(vr[jj] + gb.x);
            const float zg = sigmoid_fast(vz[jj] + gb.y);
            const float n = tanh_fast(vi[jj] + gb.z + rg * (vh[jj] + gb.w));
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
      fence_async_smem();
      fence_before_sync();
      if (mode == 0 || mode == 4) mbar_arrive_cluster(&S.a_ready, 0);     // (mode 1: the next step's x_full covers h')
      }
      EP_STAMP(8);
      if (kFuseEnv && early_x && t + 1 < T) {
        // The next step's observation tile, one step ahead: the tile (the first 8 KB of b0) is free -- the input products,
        // its last readers, completed before E4 -- and so are the accumulators the issuer will overwrite (actor.0 / fc1
        // were read by E1 / E3, R / Z / Hn by E4; Q has its own columns).  Behind q.0 the issuer then runs the next
        // step's observation and recurrent products under E5.
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          int r_, k4_;
          x_slot(tid + i * kT2EpiThreads, 0, r_, k4_);
          const float v[4] = {x4_next[i].x, x4_next[i].y, x4_next[i].z, x4_next[i].w};
          store_split4(xhi, xlo, r_, 4 * k4_, 32, v);
        }
        fence_async_smem();
        fence_before_sync();
        mbar_arrive_cluster(&S.x_full[0], 0);
      }

      if ((mode == 0 || mode == 2) && kBigA && big_smem) {
        // ---- E2 behind E4 (kBigA): the actor head on the staged table, while the Q-head product runs
        mbar_wait(&S.t_full, t_par); t_par ^= 1u;
        actor_head_big(reinterpret_cast<const float4*>(tabs + (size_t)ub * A8), [](const float4* q_) { return *q_; });
      }
      if (mode == 0 || mode == 2) {
      // ---- E5: Q tail, outputs, selection
      if (mode == 0) { epi_wait(&S.d_ready, d_par, warp); d_par ^= 1u; }   // heads only: q.0 finished with actor.2
      else epi_bar_sync();                                                  // ... but P (written by E2's last stage) must be visible
      fence_after_sync();
      EP_STAMP(9);
      if (!kBigA) {
        // NS action slots: the reference scenario has 5 actions (2 radars x {suppress, deceive} + idle), and the tail is
        // 4 FP32 instructions per (unit, slot) -- three of eight slots were computed for nothing
        auto q_tail = [&](auto ns_tag) {
          constexpr int NS = decltype(ns_tag)::value;
          float acc[NS], pa[NS];
#pragma unroll
          for (int j = 0; j < NS; ++j) { acc[j] = 0.f; pa[j] = (j < A) ? Ps[j * kTcRows + r] : 0.f; }
          for (int c0 = 0; c0 < kT2Upt; c0 += 16) {
            float v[16];
            tmem_ld16_nowait(tl + (mode == 2 ? kT2ColHn : kT2ColQ) + (uint32_t)c0, v);
            tmem_ld_wait();
            reg_fence(v);
#pragma unroll
            for (int n = 0; n < 16; ++n) {
              const int u = ub + c0 + n;
              const float4 qc = S.c.q_c[u];
              const float4 wl = S.c.w1a[2 * u], wh = NS > 4 ? S.c.w1a[2 * u + 1] : make_float4(0.f, 0.f, 0.f, 0.f);
              const float pre = v[n] + qc.x;
              const float wa[8] = {wl.x, wl.y, wl.z, wl.w, wh.x, wh.y, wh.z, wh.w};
#pragma unroll
              for (int j = 0; j < NS; ++j) acc[j] = fmaf(qc.z, fmaxf(fmaf(pa[j], qc.y, pre + wa[j]), 0.f), acc[j]);
            }
          }
          epi_bar_sync();                 // every thread has read its Ps before red is reused
#pragma unroll
          for (int j = 0; j < NS; ++j) S.red[part][j][r] = acc[j];
          epi_bar_sync();
          if (part == 0)
#pragma unroll
            for (int j = 0; j < NS; ++j)
              if (j < A) {
                float sum = bq2;
#pragma unroll
                for (int pp = 0; pp < kT2Parts; ++pp) sum += S.red[pp][j][r];
                Qs[j * kTcRows + r] = sum;
              }
        };
        if (A <= 5) q_tail(std::integral_constant<int, 5>());
        else q_tail(std::integral_constant<int, 8>());
      } else {
        auto q_tail_big = [&](const float4* w1, auto ld) {
        // this thread's 32 units of the shared hidden product stay in registers; 8 actions per pass over them
        float prev[kT2Upt];
#pragma unroll
        for (int c0 = 0; c0 < kT2Upt; c0 += 16) {
          float v[16];
          tmem_ld16_nowait(tl + (mode == 2 ? kT2ColHn : kT2ColQ) + (uint32_t)c0, v);
          tmem_ld_wait();
          reg_fence(v);
#pragma unroll
          for (int n = 0; n < 16; ++n) prev[c0 + n] = v[n] + S.c.q_c[ub + c0 + n].x;
        }
        const int a84 = A8 >> 2;
#pragma unroll 1
        for (int g = 0; g < A8; g += 8) {
          float acc[8], pa[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { acc[j] = 0.f; pa[j] = (g + j < A) ? Ps[(g + j) * kTcRows + r] : 0.f; }
#pragma unroll
          for (int n = 0; n < kT2Upt; ++n) {
            const float4 qc = S.c.q_c[ub + n];
            const float4 wl = ld(w1 + n * a84 + (g >> 2)), wh = ld(w1 + n * a84 + (g >> 2) + 1);
            const float wa[8] = {wl.x, wl.y, wl.z, wl.w, wh.x, wh.y, wh.z, wh.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(qc.z, fmaxf(fmaf(pa[j], qc.y, prev[n] + wa[j]), 0.f), acc[j]);
          }
          epi_bar_sync();               // red is free (first group: E2's last use; later groups: the previous reduction)
#pragma unroll
          for (int j = 0; j < 8; ++j) S.red[part][j][r] = acc[j];
          epi_bar_sync();
#pragma unroll
          for (int jj = 0; jj < 8 / kT2Parts; ++jj) {
            const int j = part * (8 / kT2Parts) + jj;
            if (g + j < A) {
              float sum = bq2;
#pragma unroll
              for (int pp = 0; pp < kT2Parts; ++pp) sum += S.red[pp][j][r];
              Qs[(g + j) * kTcRows + r] = sum;
            }
          }
        }
        };
        if (big_smem) q_tail_big(reinterpret_cast<const float4*>(tabs + (size_t)(H + ub) * A8), [](const float4* q_) { return *q_; });
        else q_tail_big(reinterpret_cast<const float4*>(big_w1a + (size_t)ub * A8), [](const float4* q_) { return __ldg(q_); });
        epi_bar_sync();                 // every action's Q is staged before the selection reads them
      }
      fence_before_sync();
      // (the env warps have read the previous step's actions: all but never a wait -- they had a whole agent step)
      if (kFuseEnv && t > 0) asm volatile("bar.sync %0, %1;\n" ::"n"(kT2BarEnvDone), "n"(kT2EpiThreads + kT2EnvThreads) : "memory");
      if (part == 0 && live) {
        const size_t m = tM + row0 + r;
        float best = -INFINITY, bestm = -INFINITY;
        int bi = 0, bim = 0, n_avail = 0;
        for (int act = 0; act < A; ++act) {
          const float q = Qs[act * kTcRows + r];
          const float p = Ps[act * kTcRows + r];
          if (io.q_all) io.q_all[m * A + act] = q;
          if (io.params_all) io.params_all[m * A + act] = p;
          if (q > best) { best = q; bi = act; }
          const bool ok = (av_mask >> act) & 1ull;
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
            if (u < (io.epsilon_dev ? __ldg(io.epsilon_dev + t) : io.epsilon)) {
              if (io.rand_actions) {
                chosen = io.rand_actions[m];
              } else {
                const float u2 = philox_uniform(io.seed, kStreamRandomAction, row_id, (io.rng_step_dev ? __ldg(io.rng_step_dev) : io.rng_step) + t, 0);
                const int navl = n_avail > 0 ? n_avail : A;
                int kth = (int)(u2 * (float)navl);
                kth = kth >= navl ? navl - 1 : kth;
                chosen = 0;
                for (int act = 0, seen = 0; act < A; ++act) {
                  const bool ok = (n_avail == 0) || ((av_mask >> act) & 1ull);
                  if (ok) { if (seen == kth) { chosen = act; break; } ++seen; }
                }
              }
              chosen = chosen < 0 ? 0 : (chosen >= A ? A - 1 : chosen);
            }
          }
          io.actions[m] = chosen;
          const float pw = Ps[chosen * kTcRows + r];
          if (kFuseEnv) { S.act_s[r] = chosen; S.pow_s[r] = pw; }
          if (io.power) io.power[m] = pw;
          if (io.actions_mirror) io.actions_mirror[m] = chosen;      // e.g. the caller's page-locked host copy
          if (io.power_mirror) io.power_mirror[m] = pw;
          if (io.q_chosen) io.q_chosen[m] = Qs[chosen * kTcRows + r];
        }
      }
      }
      if (kFuseEnv) {
        // ---- the actions just chosen go to the env warps, which run the step of this CTA's envs from here on
        EP_STAMP(11);
        __threadfence_block();
        asm volatile("bar.arrive %0, %1;\n" ::"n"(kT2BarActions), "n"(kT2EpiThreads + kT2EnvThreads) : "memory");
      }
      epi_bar_sync();                   // Ps / Qs / red are free for the next step
      EP_STAMP(10);
    }
  }
  fence_before_sync();
  __syncthreads();
  TC_STAMP_ONCE(24);
  cluster_sync_all();                   // the leader's MMAs read the peer's shared memory until here
  if (warp == kT2EpiWarps + 1) tmem_dealloc_2sm(tmem, kT2TmemCols);
  TC_STAMP_ONCE(25);
  TC_CTA_STAMP(1);
}

inline size_t agent_tc2_smem_bytes(const macjd_agent_weights& w) {
  // Ps [A][64] and, up to 8 actions, Qs [A][64] behind it (kBigA: Qs borrows the xf tile)
  return sizeof(T2Smem) + sizeof(float) * (w.n_actions > 8 ? 1 : 2) * (size_t)w.n_actions * kTcRows + 1024;
}

inline bool agent_tc2_supported(const macjd_agent_weights& w) {
  return agent_tc_supported(w) && w.tc_format == 1 && kTcKc == 32 && kTcChunksPerX == 1 && agent_tc2_smem_bytes(w) <= 227 * 1024;
}

// The fused rollout step needs whole envs per CTA (64 % J == 0), derived scenario tables and whole warps of lane groups
// (one group of G lanes per env; the env warps take them in passes of kT2EnvThreads lanes).
inline bool agent_tc2_fuse_supported(const macjd_agent_weights& w, const macjd_env_tables& t) {
  if (!agent_tc2_supported(w) || t.derived == nullptr || t.n_jammers < 1 || kTcRows % t.n_jammers != 0 || !env2_supported(t)) return false;
  const int epc = kTcRows / t.n_jammers, G = env2_group(t.n_jammers, t.n_radars);
  return epc * G <= 2 * kT2EnvThreads && (epc * G) % 32 == 0;
}

// ... and it pays off while the env work a CTA takes on is small next to its agent step: the views a CTA copies per
// step (state + J observation rows + availability of 64 / J envs) stay within 16 KB.  Measured on B200: default
// scenario (9 KB per CTA-step) 37.1 -> 35.5 us per step fused; 8 x 16 x 4 scenario (50 KB) 521 us fused against
// the two kernels overlapping (MACJD_FUSE_MAX_VIEW_BYTES overrides the limit, for experiments).
inline bool agent_tc2_fuse_profitable(const macjd_env_tables& t) {
  const char* ev = getenv("MACJD_FUSE_MAX_VIEW_BYTES");            // read per call (tens of ns): tests switch it
  const size_t limit = ev ? (size_t)strtoull(ev, nullptr, 10) : (size_t)16384;
  const size_t S = (size_t)t.n_radars * (6 + t.n_types) + 2 * (size_t)t.n_jammers, J = (size_t)t.n_jammers;
  const size_t per_env = S * 4 * (1 + J) + J * (2 * (size_t)t.n_radars + 1);
  return per_env * (kTcRows / J) <= limit;
}

// TMA descriptor of a packed chunk buffer, seen as [bytes / 128][32 floats]: one box of 64 rows is one contiguous 8 KB
// half-chunk (a pure host-side encode; the driver entry point is looked up once)
inline bool encode_chunk_map(CUtensorMap* map, const float* chunks, size_t chunk_bytes) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      fn = nullptr;
    return reinterpret_cast<EncodeFn>(fn);
  }();
  if (!encode) return false;
  const cuuint64_t gdim[2] = {32, chunk_bytes / 128};
  const cuuint64_t gstride[1] = {128};
  const cuuint32_t box[2] = {32, kT2HalfBytes / 128};
  const cuuint32_t estride[2] = {1, 1};
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(chunks), gdim, gstride, box, estride,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline int agent_tc2_launch(const macjd_ctx* ctx, const AgentArgs& a, const Env2Args* env = nullptr) {
  const size_t smem = agent_tc2_smem_bytes(a.w);
  // the opt-in is per device and sticky: ask once per device and size (an act call is latency-critical)
  static PerDeviceMax opted;
  if (!opted.covers(ctx->device, smem)) {
    if (cudaFuncSetAttribute(agent_forward_tc2_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(agent_forward_tc2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(agent_forward_tc2_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(agent_forward_tc2_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(agent_forward_tc2_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(agent_forward_tc2_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return MACJD_ERR_CUDA;
    opted.record(ctx->device, smem);
  }
  if (watchdog_arm(ctx->device) != MACJD_OK) return MACJD_ERR_CUDA;
  T2Args p;
  p.a = a;
  if (env) p.env = *env;
  const size_t chunk_bytes = (size_t)(2 * kTcChunksPerX * (a.w.obs_pad / 32) + 8 * kTcChunksPerH) * kTcChunkBytes;
  if (!encode_chunk_map(&p.wmap, a.w.tc_chunks, chunk_bytes)) return MACJD_ERR_CUDA;
  const int pairs = (a.io.n_rows + 2 * kTcRows - 1) / (2 * kTcRows);
  const bool big = a.w.n_actions > 8;
  const cudaStream_t st = (cudaStream_t)ctx->stream;
  if (env) {
    if (a.io.part != 0 || a.io.n_steps < 1) return MACJD_ERR_INVALID_ARG;
    if (big) agent_forward_tc2_kernel<true, true, true><<<2 * pairs, kT2Threads + kT2EnvThreads, smem, st>>>(p);
    else agent_forward_tc2_kernel<true, false, true><<<2 * pairs, kT2Threads + kT2EnvThreads, smem, st>>>(p);
  } else if (a.io.part == 0) {
    if (big) agent_forward_tc2_kernel<true, true><<<2 * pairs, kT2Threads, smem, st>>>(p);
    else agent_forward_tc2_kernel<true, false><<<2 * pairs, kT2Threads, smem, st>>>(p);
  } else {
    if (big) agent_forward_tc2_kernel<false, true><<<2 * pairs, kT2Threads, smem, st>>>(p);
    else agent_forward_tc2_kernel<false, false><<<2 * pairs, kT2Threads, smem, st>>>(p);
  }
  return MACJD_OK;
}

#if defined(MACJD_TC_PROFILE) || defined(MACJD_DEBUG_TOOLS)   // tooling build only (tools/tc_mma_rate.py)
// Micro-benchmark: like tc_mma_rate_kernel, for tcgen05.mma.cta_group::2 issued by the leader of a CTA pair.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) tc2_mma_rate_kernel(int M, int N, int n, unsigned long long* out) {
  extern __shared__ __align__(128) unsigned char tc_smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  for (int i = tid; i < 40 * 1024 / 4; i += 128) reinterpret_cast<float*>(tc_smem)[i] = 0.001f * (float)(i & 255);
  fence_async_smem();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_alloc_2sm(&tmem_base_s, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  if (warp == 0 && rank == 0) {
    // whole warp converged, one elected lane issues: operands stay in uniform registers (issued from
    // inside `if (tid == 0)` every MMA is wrapped in an ELECT / R2UR sequence that costs ~50 cycles)
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t idesc = umma_idesc_tf32(M, N);
    const uint64_t da = umma_smem_desc(smem_u32(tc_smem), 128, 16 * 32);
    const uint64_t db = umma_smem_desc(smem_u32(tc_smem) + 16384, 128, 16 * 32);
    const unsigned long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
      const uint32_t d = tb + (uint32_t)((i & 1) * 256);
      const uint64_t a = da + (uint64_t)((i & 1) * 16);
      if (elect_one()) mma_tf32_ss_2sm(d, a, db, idesc, 1u);
      __syncwarp();
    }
    if (elect_one()) mma_commit_2sm(&bar);
    __syncwarp();
    const unsigned long long t1 = clock64();
    mbar_wait_cluster(&bar, 0);
    const unsigned long long t2 = clock64();
    if (tid == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  } else if (tid == 0) {
    mbar_wait_cluster(&bar, 0);
  }
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_2sm(tmem_base, 512);
}

inline int tc2_mma_rate(const macjd_ctx* ctx, int M, int N, int n, unsigned long long* out_dev) {
  if (cudaFuncSetAttribute(tc2_mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024) != cudaSuccess)
    return MACJD_ERR_CUDA;
  tc2_mma_rate_kernel<<<2, 128, 64 * 1024, (cudaStream_t)ctx->stream>>>(M, N, n, out_dev);
  return MACJD_OK;
}

#endif  // tooling build

}  // namespace tc
}  // namespace macjd
#endif  // !MACJD_TEST_HOST_EMULATION
