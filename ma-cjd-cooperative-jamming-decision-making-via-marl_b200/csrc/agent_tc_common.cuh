// What the tcgen05 agent kernel (agent_act_tc2.cuh) shares with its tooling: chunk geometry of the
// packed weights, mbarrier / bulk-copy helpers, the operand-tile split store, optional phase stamps.
// (Round 1 also had a single-CTA kernel, M = 64 per MMA, here; the CTA-pair kernel superseded it for
// every shape -- 22 vs 33 us per step even at <= 64 rows -- and it was removed from the product.)
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
constexpr int kTcChunksPerH = kTcH / kTcKc;               // chunks of a K = 128 layer
constexpr int kTcChunksPerX = 32 / kTcKc;                 // chunks per 32-wide observation block
constexpr uint32_t kTcAStep = kTcKc * 32;                 // A-tile byte offset between chunks
constexpr int kTcChunkFloats = 2 * kTcH * kTcKc;          // hi + lo
constexpr int kTcChunkBytes = kTcChunkFloats * 4;         // 16 KB

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
#ifdef MACJD_TC_PROFILE
#define EP_STAMP(slot) do { if (tid == 0) TC_STAMP(slot); } while (0)
#else
#define EP_STAMP(slot) do { } while (0)
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

inline int agent_tc_chunk_k() { return kTcKc; }

// dims the packed tensor-core weights exist for (core/networks.py: PackedAgentWeights._pack_tc)
inline bool agent_tc_supported(const macjd_agent_weights& w) {
  return w.tc_chunks != nullptr && w.hidden == kTcH && w.actor_hidden == kTcH && w.obs_pad % 32 == 0 && w.n_actions <= 64;
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
