// Episode replay ring: strided, index-driven episode copies (K11 of SURVEY.md).
//
// One kernel serves both directions of utils/replay_buffer.py:
//   store  (replay_buffer.py:78-151): rollout buffers (time-major) -> ring slots   index on dst
//   sample (replay_buffer.py:153-214): ring slots -> training batch (any layout)  index on src
// For every key the host passes one descriptor: base pointers, bytes per timestep, number of
// timesteps and the (episode, timestep) byte strides on both sides.  Pure data movement,
// HBM-bound: 2 x 142 792 B per episode at the default configuration.
#pragma once
#include "macjd_common.cuh"

namespace macjd {

struct CopyArgs {
  macjd_copy_desc key[MACJD_MAX_COPY_KEYS];
  int n_keys;
  const int32_t* idx;   // [n_eps] ring slot per batch entry, or null (identity)
  int n_eps;
  int index_on_src;
};

template <typename V>
__device__ __forceinline__ void copy_rows(const macjd_copy_desc& k, int64_t s_ep, int64_t d_ep, int tid, int nthreads) {
  const int inner = k.inner_bytes / (int)sizeof(V);
  const int total = k.n_t * inner;
  const char* src = reinterpret_cast<const char*>(k.src) + s_ep * k.src_ep_stride;
  char* dst = reinterpret_cast<char*>(k.dst) + d_ep * k.dst_ep_stride;
  for (int e = tid; e < total; e += nthreads) {
    const int t = e / inner, v = e - t * inner;
    const V val = *reinterpret_cast<const V*>(src + (int64_t)t * k.src_t_stride + (int64_t)v * sizeof(V));
    *reinterpret_cast<V*>(dst + (int64_t)t * k.dst_t_stride + (int64_t)v * sizeof(V)) = val;
  }
}

// float32 <-> bfloat16 (round to nearest even; NaN stays NaN), bit arithmetic so that host emulation runs the same code
__host__ __device__ __forceinline__ uint16_t f32_to_bf16_bits(uint32_t u) {
  if ((u & 0x7F800000u) == 0x7F800000u && (u & 0x007FFFFFu)) return (uint16_t)((u >> 16) | 0x0040u);   // NaN
  return (uint16_t)((u + 0x7FFFu + ((u >> 16) & 1u)) >> 16);
}
// convert == 1: float32 -> bfloat16, convert == 2: bfloat16 -> float32; k.inner_bytes counts SOURCE bytes.
// vec4: four elements per access (16-byte float4 / 8-byte bf16 x 4), else one.
template <bool kVec4>
__device__ __forceinline__ void convert_rows(const macjd_copy_desc& k, int64_t s_ep, int64_t d_ep, int tid, int nthreads) {
  const bool to_bf16 = k.convert == 1;
  const int elems = k.inner_bytes / (to_bf16 ? 4 : 2), per = kVec4 ? 4 : 1;
  const int inner = elems / per, total = k.n_t * inner;
  const char* src = reinterpret_cast<const char*>(k.src) + s_ep * k.src_ep_stride;
  char* dst = reinterpret_cast<char*>(k.dst) + d_ep * k.dst_ep_stride;
  for (int e = tid; e < total; e += nthreads) {
    const int t = e / inner, v = (e - t * inner) * per;
    const char* s = src + (int64_t)t * k.src_t_stride + (int64_t)v * (to_bf16 ? 4 : 2);
    char* d = dst + (int64_t)t * k.dst_t_stride + (int64_t)v * (to_bf16 ? 2 : 4);
    if (to_bf16) {
      if (kVec4) {
        const uint4 x = *reinterpret_cast<const uint4*>(s);
        uint2 y;
        y.x = (uint32_t)f32_to_bf16_bits(x.x) | ((uint32_t)f32_to_bf16_bits(x.y) << 16);
        y.y = (uint32_t)f32_to_bf16_bits(x.z) | ((uint32_t)f32_to_bf16_bits(x.w) << 16);
        *reinterpret_cast<uint2*>(d) = y;
      } else {
        *reinterpret_cast<uint16_t*>(d) = f32_to_bf16_bits(*reinterpret_cast<const uint32_t*>(s));
      }
    } else {
      if (kVec4) {
        const uint2 x = *reinterpret_cast<const uint2*>(s);
        *reinterpret_cast<uint4*>(d) = make_uint4(x.x << 16, x.x & 0xFFFF0000u, x.y << 16, x.y & 0xFFFF0000u);
      } else {
        *reinterpret_cast<uint32_t*>(d) = (uint32_t)(*reinterpret_cast<const uint16_t*>(s)) << 16;
      }
    }
  }
}

// grid = (n_eps, n_keys); one CTA copies one key of one episode
__global__ void __launch_bounds__(256) replay_copy_kernel(const CopyArgs a) {
  grid_dependency_sync();   // no-op unless launched as a programmatic dependent (MACJD_LAUNCH)
  const int b = blockIdx.x;
  const macjd_copy_desc& k = a.key[blockIdx.y];
  const int64_t slot = a.idx ? a.idx[b] : b;
  const int64_t s_ep = a.index_on_src ? slot : b;
  const int64_t d_ep = a.index_on_src ? b : slot;
  if (k.convert) {
    if (k.vec_bytes == 16) convert_rows<true>(k, s_ep, d_ep, threadIdx.x, blockDim.x);
    else convert_rows<false>(k, s_ep, d_ep, threadIdx.x, blockDim.x);
    return;
  }
  if (k.vec_bytes == 16) copy_rows<uint4>(k, s_ep, d_ep, threadIdx.x, blockDim.x);
  else if (k.vec_bytes == 4) copy_rows<uint32_t>(k, s_ep, d_ep, threadIdx.x, blockDim.x);
  else copy_rows<uint8_t>(k, s_ep, d_ep, threadIdx.x, blockDim.x);
}

inline int replay_copy(const macjd_ctx* ctx, const macjd_copy_desc* descs, int n_keys, const int32_t* idx, int n_eps,
                       int index_on_src) {
  if (!ctx || !descs || n_keys < 1 || n_keys > MACJD_MAX_COPY_KEYS || n_eps < 0) return MACJD_ERR_INVALID_ARG;
  if (n_eps == 0) return MACJD_OK;
  CopyArgs a;
  a.n_keys = n_keys; a.idx = idx; a.n_eps = n_eps; a.index_on_src = index_on_src;
  for (int i = 0; i < n_keys; ++i) {
    macjd_copy_desc d = descs[i];
    if (!d.src || !d.dst || d.n_t < 0 || d.inner_bytes < 0 || d.convert < 0 || d.convert > 2) return MACJD_ERR_INVALID_ARG;
    if (d.convert) {
      // four elements per access when the float32 side is 16-byte and the bfloat16 side 8-byte aligned throughout
      const int fb = d.convert == 1 ? 4 : 2;                       // source element bytes
      if (d.inner_bytes % fb) return MACJD_ERR_INVALID_ARG;
      const uint64_t sb = (uint64_t)(uintptr_t)d.src | (uint64_t)d.src_ep_stride | (uint64_t)d.src_t_stride;
      const uint64_t db = (uint64_t)(uintptr_t)d.dst | (uint64_t)d.dst_ep_stride | (uint64_t)d.dst_t_stride;
      const uint64_t f32_side = d.convert == 1 ? sb : db, bf_side = d.convert == 1 ? db : sb;
      const bool v4 = (d.inner_bytes / fb) % 4 == 0 && f32_side % 16 == 0 && bf_side % 8 == 0;
      d.vec_bytes = v4 ? 16 : 1;
      a.key[i] = d;
      continue;
    }
    // widest vector that keeps every access aligned
    const uint64_t bits = (uint64_t)(uintptr_t)d.src | (uint64_t)(uintptr_t)d.dst | (uint64_t)d.inner_bytes |
                          (uint64_t)d.src_ep_stride | (uint64_t)d.src_t_stride | (uint64_t)d.dst_ep_stride |
                          (uint64_t)d.dst_t_stride;
    d.vec_bytes = (bits % 16 == 0) ? 16 : (bits % 4 == 0) ? 4 : 1;
    a.key[i] = d;
  }
  MACJD_LAUNCH(replay_copy_kernel, dim3(n_eps, n_keys), dim3(256), 0, (cudaStream_t)ctx->stream, a);
  return MACJD_OK;
}

}  // namespace macjd
