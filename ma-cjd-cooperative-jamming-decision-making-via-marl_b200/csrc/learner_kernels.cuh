// Row-wise kernels of the QMix learner (K5/K7/K8/K9 of SURVEY.md): LayerNorm, the mixing
// network proper (clamped hypernet outputs -> ELU layer -> Q_tot) forward and backward, the
// MP-DQN Q-head tail on stored hidden states, the double-DQN TD target / masked loss, column
// reductions for bias / LayerNorm gradients, and gradient-norm + clip + Adam.
// Reference: core/networks.py:250-316 (QMixer.forward), core/networks.py:131-180,
// core/qmix.py:138-200.  One warp per row; lanes run along the feature dimension.
#pragma once
#include "macjd_common.cuh"

namespace macjd {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------- LayerNorm (networks.py:215,270)
// y = (x - mean) * rstd * gamma + beta, biased variance, eps = 1e-5.  Saves xhat for dgamma.
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, int R, int S,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ y,
                                                            float* __restrict__ xhat) {
  grid_dependency_sync();   // no-op unless launched as a programmatic dependent (MACJD_LAUNCH)
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;   // whole warps exit together; no block-level sync below
  const float* xr = x + (size_t)row * S;
  float s = 0.f;
  for (int i = lane; i < S; i += 32) s += xr[i];
  const float mean = warp_sum(s) / (float)S;
  float v = 0.f;
  for (int i = lane; i < S; i += 32) { const float d = xr[i] - mean; v = fmaf(d, d, v); }
  const float rstd = 1.0f / sqrtf(warp_sum(v) / (float)S + 1e-5f);
  for (int i = lane; i < S; i += 32) {
    const float h = (xr[i] - mean) * rstd;
    if (xhat) xhat[(size_t)row * S + i] = h;
    y[(size_t)row * S + i] = fmaf(h, gamma[i], beta[i]);
  }
}

// ---------------------------------------------------------------- mixing network (networks.py:301-307)
// hidden = elu(q W1 + b1);  y = hidden . wf + v       (W1 [N][E], b1 [E], wf [E], v: per row,
// already clamped by the producing GEMM epilogues)
__global__ void __launch_bounds__(256) mix_fwd_kernel(const float* __restrict__ q, const float* __restrict__ w1,
                                                      const float* __restrict__ b1, const float* __restrict__ wf,
                                                      const float* __restrict__ v, int R, int N, int E,
                                                      float* __restrict__ hidden, float* __restrict__ y) {
  grid_dependency_sync();   // no-op unless launched as a programmatic dependent (MACJD_LAUNCH)
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  const float* qr = q + (size_t)row * N;
  const float* w1r = w1 + (size_t)row * N * E;
  float acc = 0.f;
  for (int e = lane; e < E; e += 32) {
    float pre = b1[(size_t)row * E + e];
    for (int n = 0; n < N; ++n) pre = fmaf(qr[n], w1r[(size_t)n * E + e], pre);
    const float h = pre > 0.f ? pre : expm1f(pre);          // F.elu, alpha = 1
    if (hidden) hidden[(size_t)row * E + e] = h;
    acc = fmaf(h, wf[(size_t)row * E + e], acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) y[row] = acc + v[row];
}

// Backward of the mixing network for one row given dy = dL/dQ_tot.  Produces the gradients
// w.r.t. the *raw* (pre-clamp) hypernet outputs -- clamp passes gradient strictly inside its
// range -- and dq = dL/dq_i.
__global__ void __launch_bounds__(256) mix_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ q,
                                                      const float* __restrict__ w1, const float* __restrict__ b1,
                                                      const float* __restrict__ wf, const float* __restrict__ v,
                                                      const float* __restrict__ hidden, int R, int N, int E,
                                                      float* __restrict__ d_w1, float* __restrict__ d_b1,
                                                      float* __restrict__ d_wf, float* __restrict__ d_v,
                                                      float* __restrict__ dq) {
  grid_dependency_sync();   // no-op unless launched as a programmatic dependent (MACJD_LAUNCH)
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  const float g = dy[row];
  const float* qr = q + (size_t)row * N;
  const float* w1r = w1 + (size_t)row * N * E;
  if (lane == 0) { const float c = v[row]; d_v[row] = (c > -5.f && c < 5.f) ? g : 0.f; }
  for (int n = 0; n < N; ++n) {
    float part = 0.f;
    for (int e = lane; e < E; e += 32) {
      const float h = hidden[(size_t)row * E + e];
      const float cwf = wf[(size_t)row * E + e];
      const float dpre = g * cwf * (h > 0.f ? 1.f : h + 1.f);   // elu' = 1 or exp(pre) = h + 1
      const float cw1 = w1r[(size_t)n * E + e];
      part = fmaf(cw1, dpre, part);
      d_w1[((size_t)row * N + n) * E + e] = (cw1 > 0.f && cw1 < 5.f) ? qr[n] * dpre : 0.f;
      if (n == 0) {
        const float cb1 = b1[(size_t)row * E + e];
        d_b1[(size_t)row * E + e] = (cb1 > -5.f && cb1 < 5.f) ? dpre : 0.f;
        d_wf[(size_t)row * E + e] = (cwf > 0.f && cwf < 5.f) ? g * h : 0.f;
      }
    }
    part = warp_sum(part);
    if (lane == 0 && dq) dq[(size_t)row * N + n] = part;
  }
}

// ---------------------------------------------------------------- Q-head tail (networks.py:147-180)
// hid = relu(pre + W1[:, H+a] + p W1[:, H+A]);  q = w2 . hid + b2      (pre = h W1[:, :H]^T + b1)
__global__ void __launch_bounds__(256) qhead_tail_fwd_kernel(float* __restrict__ hid, const int* __restrict__ act,
                                                             const float* __restrict__ par,
                                                             const float* __restrict__ w1a,
                                                             const float* __restrict__ w1p,
                                                             const float* __restrict__ w2,
                                                             const float* __restrict__ b2, int R, int H, int A,
                                                             float* __restrict__ q) {
  grid_dependency_sync();   // no-op unless launched as a programmatic dependent (MACJD_LAUNCH)
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  int a = act[row];
  a = a < 0 ? 0 : (a >= A ? A - 1 : a);
  const float p = par[row];
  float acc = 0.f;
  for (int n = lane; n < H; n += 32) {
    float h = hid[(size_t)row * H + n] + w1a[(size_t)a * H + n] + p * w1p[n];
    h = fmaxf(h, 0.f);
    hid[(size_t)row * H + n] = h;
    acc = fmaf(h, w2[n], acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) q[row] = acc + b2[0];
}

// dhid = dq * w2 * 1[hid > 0]  (in place over hid), xaug = [onehot(a), p]
__global__ void __launch_bounds__(256) qhead_tail_bwd_kernel(float* __restrict__ hid, const float* __restrict__ dq,
                                                             const int* __restrict__ act,
                                                             const float* __restrict__ par,
                                                             const float* __restrict__ w2, int R, int H, int A,
                                                             float* __restrict__ xaug) {
  grid_dependency_sync();   // no-op unless launched as a programmatic dependent (MACJD_LAUNCH)
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  const float g = dq[row];
  for (int n = lane; n < H; n += 32) {
    const float h = hid[(size_t)row * H + n];
    hid[(size_t)row * H + n] = h > 0.f ? g * w2[n] : 0.f;
  }
  int a = act[row];
  a = a < 0 ? 0 : (a >= A ? A - 1 : a);
  for (int c = lane; c <= A; c += 32) xaug[(size_t)row * (A + 1) + c] = c == A ? par[row] : (c == a ? 1.f : 0.f);
}

// ---------------------------------------------------------------- double-DQN gather (qmix.py:147)
// out[i] = q_all[i][idx[i]]   (i over rows x timesteps, q_all [.., A])
__global__ void __launch_bounds__(256) gather_q_kernel(const float* __restrict__ q_all, const int* __restrict__ idx,
                                                       int n, int A, float* __restrict__ out) {
  grid_dependency_sync();   // no-op unless launched as a programmatic dependent (MACJD_LAUNCH)
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int a = idx[i];
  a = a < 0 ? 0 : (a >= A ? A - 1 : a);
  out[i] = q_all[(size_t)i * A + a];
}

// ---------------------------------------------------------------- column reductions
// part[chunk][c] = sum over the chunk's rows of X[r][c] (* Y[r][c]);  then a fixed-order
// second pass.  Used for bias gradients and LayerNorm dgamma / dbeta.
constexpr int kColsumRows = 64;     // rows per partial block: each thread walks them serially, so short chunks, many blocks
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ X, const float* __restrict__ Y,
                                                             int R, int Cn, int ldx, float* __restrict__ part, int rows_per_chunk) {
  grid_dependency_sync();   // no-op unless launched as a programmatic dependent (MACJD_LAUNCH)
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Cn) return;
  const int r0 = blockIdx.y * rows_per_chunk, r1 = min(R, r0 + rows_per_chunk);
  // eight rows in flight per thread (independent partial sums, combined in a fixed order): with one running sum the loop
  // was a chain of dependent loads -- 70 us per 100 MB matrix at 101 376 rows (ncu launch list of a C4 train step)
  float s8[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) s8[u] = 0.f;
  int r = r0;
  for (; r + 8 <= r1; r += 8) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float x = X[(size_t)(r + u) * ldx + c];
      s8[u] += Y ? x * Y[(size_t)(r + u) * ldx + c] : x;
    }
  }
  for (; r < r1; ++r) {
    const float x = X[(size_t)r * ldx + c];
    s8[0] += Y ? x * Y[(size_t)r * ldx + c] : x;
  }
  part[(size_t)blockIdx.y * Cn + c] = ((s8[0] + s8[1]) + (s8[2] + s8[3])) + ((s8[4] + s8[5]) + (s8[6] + s8[7]));
}
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ part, int chunks, int Cn,
                                                           float* __restrict__ out) {
  grid_dependency_sync();   // no-op unless launched as a programmatic dependent (MACJD_LAUNCH)
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Cn) return;
  float s = 0.f;
  for (int k = 0; k < chunks; ++k) s += part[(size_t)k * Cn + c];
  out[c] = s;
}
inline size_t colsum_ws_floats(int R, int Cn) { return (size_t)((R + kColsumRows - 1) / kColsumRows) * Cn; }
inline void colsum(cudaStream_t st, const float* X, const float* Y, int R, int Cn, int ldx, float* out, float* ws) {
  if (Cn <= 0) return;
  // at most 256 chunks (the second pass walks them serially in one block per 256 columns: 1 584 chunks of 64 rows at
  // 101 376 rows made it 93 us), at least 64 rows each
  int rows = (R + 255) / 256;
  if (rows < kColsumRows) rows = kColsumRows;
  const int chunks = (R + rows - 1) / rows;
  if (chunks == 0) { cudaMemsetAsync(out, 0, sizeof(float) * Cn, st); return; }
  MACJD_LAUNCH(colsum_partial_kernel, dim3((Cn + 255) / 256, chunks), dim3(256), 0, st, X, Y, R, Cn, ldx, ws, rows);
  MACJD_LAUNCH(colsum_final_kernel, dim3((Cn + 255) / 256), dim3(256), 0, st, (const float*)ws, chunks, Cn, out);
}

// ---------------------------------------------------------------- TD target / loss (qmix.py:155,191-194)
// rows r = t*B + b, t = 0..T-2.   targets = reward + gamma (1 - terminated) tq_tot
// td = q_tot - targets;  dy = 2 td mask (un-normalised; 1/sum(mask) is applied with the clip)
// sums[0..3] += { sum (td mask)^2, sum mask, sum q_tot, sum targets }  via per-block partials.
constexpr int kTdBlock = 256;
__global__ void __launch_bounds__(kTdBlock) td_partial_kernel(const float* __restrict__ q_tot,
                                                              const float* __restrict__ tq_tot,
                                                              const float* __restrict__ reward,
                                                              const uint8_t* __restrict__ terminated,
                                                              const uint8_t* __restrict__ filled, float gamma, int R,
                                                              float* __restrict__ dy, float* __restrict__ targets_out,
                                                              float* __restrict__ part) {
  grid_dependency_sync();   // no-op unless launched as a programmatic dependent (MACJD_LAUNCH)
  __shared__ float red[4][kTdBlock / 32];
  const int r = blockIdx.x * kTdBlock + threadIdx.x;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (r < R) {
    const float m = filled[r] ? 1.f : 0.f;
    const float tgt = reward[r] + gamma * (1.f - (terminated[r] ? 1.f : 0.f)) * tq_tot[r];
    const float td = (q_tot[r] - tgt) * m;
    dy[r] = 2.f * td * m;
    if (targets_out) targets_out[r] = tgt;
    s0 = td * td; s1 = m; s2 = q_tot[r]; s3 = tgt;
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2); s3 = warp_sum(s3);
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][w] = s0; red[1][w] = s1; red[2][w] = s2; red[3][w] = s3; }
  __syncthreads();
  if (threadIdx.x < 4) {
    float s = 0.f;
    for (int k = 0; k < kTdBlock / 32; ++k) s += red[threadIdx.x][k];
    part[(size_t)blockIdx.x * 4 + threadIdx.x] = s;
  }
}
__global__ void td_final_kernel(const float* __restrict__ part, int blocks, float* __restrict__ sums) {
  grid_dependency_sync();   // no-op unless launched as a programmatic dependent (MACJD_LAUNCH)
  if (threadIdx.x < 4) {
    float s = 0.f;
    for (int k = 0; k < blocks; ++k) s += part[(size_t)k * 4 + threadIdx.x];
    sums[threadIdx.x] = s;
  }
}

// ---------------------------------------------------------------- grad norm, clip, Adam (qmix.py:197-200)
constexpr int kMaxOptTensors = 32;
struct OptTable {
  float* param[kMaxOptTensors];
  int offset[kMaxOptTensors + 1];   // element offsets into the flat grad / m / v buffers
  int count;
};

__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, int n, float* __restrict__ part) {
  grid_dependency_sync();   // no-op unless launched as a programmatic dependent (MACJD_LAUNCH)
  __shared__ float red[8];
  float s = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s = fmaf(g[i], g[i], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k];
    part[blockIdx.x] = t;
  }
}

// scal[0] = grad_norm (of the normalised gradient), scal[1] = clip coefficient * scale
// sums[1] = sum(mask) (possibly all-reduced); scale = 1 / sum(mask)
__global__ void clip_coef_kernel(const float* __restrict__ part, int blocks, const float* __restrict__ sums,
                                 float max_norm, float* __restrict__ scal) {
  grid_dependency_sync();   // no-op unless launched as a programmatic dependent (MACJD_LAUNCH)
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int k = 0; k < blocks; ++k) s += (double)part[k];
    const float scale = 1.0f / sums[1];
    const float norm = (float)sqrt(s) * scale;
    float coef = max_norm / (norm + 1e-6f);          // torch.nn.utils.clip_grad_norm_
    coef = coef > 1.f ? 1.f : coef;
    scal[0] = norm;
    scal[1] = coef * scale;
    scal[2] = sums[0] * scale;                        // loss = sum (td mask)^2 / sum mask
  }
}

__global__ void __launch_bounds__(256) adam_kernel(OptTable tab, const float* __restrict__ grad, float* __restrict__ m,
                                                   float* __restrict__ v, const float* __restrict__ scal, float lr,
                                                   float beta1, float beta2, float eps, float bc1, float bc2_sqrt,
                                                   const float* __restrict__ bias_corr) {
  grid_dependency_sync();   // no-op unless launched as a programmatic dependent (MACJD_LAUNCH)
  // a captured launch (CUDA graph) serves every step: the step's bias corrections then come from device memory
  if (bias_corr) { bc1 = bias_corr[0]; bc2_sqrt = bias_corr[1]; }
  const int total = tab.offset[tab.count];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int k = 0;
  while (k + 1 < tab.count && i >= tab.offset[k + 1]) ++k;
  const float g = grad[i] * scal[1];
  const float mi = m[i] + (g - m[i]) * (1.f - beta1);          // exp_avg.lerp_(grad, 1 - beta1)
  const float vi = v[i] * beta2 + (1.f - beta2) * g * g;       // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  float* p = tab.param[k] + (i - tab.offset[k]);
  *p = *p - (lr / bc1) * (mi / denom);
}

}  // namespace macjd
