// Batched environment step on DERIVED scenario tables (round 2; the kernel the Python environment runs).
//
// Everything of simulation/environment.py:221-477 that depends on the scenario alone is evaluated ONCE per
// scenario by env_prepare_kernel (the reference rebuilds its Radar / Jammer objects every episode and
// re-derives these numbers every step):
//   per radar-target pair   echo signal Ga Ps (core/radar.py:35-60), SNR without jamming, its Albersheim Pd
//                           (core/radar.py:67-82) -- which is also the step's Pd for every radar that no
//                           suppression jammer points at (jam = 0: the same division, bit for bit)
//   per radar               noise power, D, receive gain, the clipped tracking penalty (environment.py:365)
//   per jammer-radar link   the Friis denominator d^2 L L_atm B_j (core/jammer.py:73-98)
//   per env                 the float32 state row (environment.py:479-510), row-major
// stored as one contiguous block of doubles PER ENV (array of structures: a group of lanes working on one env reads
// neighbouring addresses).  The step itself then costs one division per acting jammer, one Albersheim evaluation
// per deception attempt and per target of a SUPPRESSED radar, and the Bernoulli draws.
//
// Mapping (north_star: "warp-shuffle reductions over radars and jammers"): a GROUP of G = 2^k >= max(J, R) lanes
// steps one env -- lane j evaluates jammer j, then lane r evaluates radar r and its K targets; the jammer records
// travel to the radar lanes by shuffles and are applied in jammer order, the per-radar terms are summed by
// shuffles in radar order, so every sum and product has the reference's (sequential) order and value.
// The stand-alone kernel runs 128 physics threads (128 / G envs, at most 32) and 128 view threads per block: the
// view warps copy the static views of the same envs (state, obs = the state once per jammer, all-ones availability:
// 6.3 KB per env-step at 8 x 16 x 4, the bulk of the kernel's traffic) with 16-byte loads and stores.  Scenario
// values that do not depend on the actions are fetched before the dependency wait (programmatic dependent launch
// behind the agent kernel).  The CTA-pair agent kernel calls the same device functions when it runs the env step
// of its own rows' envs (agent_act_tc2.cuh, kFuseEnv).
#pragma once
#include <stdlib.h>
#include "env_step.cuh"

namespace macjd {

// offsets (in doubles) inside one env's derived block
struct DerivedRows {
  int pd0, sig, snr0;      // [R*K] each, slot = r * K + k
  int rad;                 // [R][4]: pn, D, clipped tracking penalty, receive gain
  int den;                 // [J][R] link denominators (-1: jammer closer than 1e-6 to the radar, no effect)
  int jam;                 // [J][4]: gj, power_min, power_max, 0
  int total;               // doubles per env (a multiple of 2: blocks stay 16-byte aligned)
};
__host__ __device__ inline DerivedRows derived_rows(int J, int R, int K) {
  DerivedRows d;
  int o = 0;
  d.pd0 = o; o += R * K; d.sig = o; o += R * K;
  d.rad = o; o += 4 * R;
  d.den = o; o += J * R;
  d.jam = o; o += 4 * J;
  d.snr0 = o; o += R * K;          // last: only the optional snr0 / snr1 outputs read it (a step without them never fetches it)
  d.total = (o + 1) & ~1;
  return d;
}
inline size_t env_derived_bytes(const macjd_env_tables& t) {
  const size_t cols = t.env_stride == 0 ? 1 : (size_t)t.n_envs;
  const size_t S = (size_t)t.n_radars * (6 + t.n_types) + 2 * (size_t)t.n_jammers;
  const size_t dbl = (size_t)derived_rows(t.n_jammers, t.n_radars, t.n_targets).total * cols * sizeof(double);
  return ((dbl + 15) & ~(size_t)15) + ((cols * S * sizeof(float) + 15) & ~(size_t)15);
}
// the float32 state rows sit behind the double blocks
__host__ __device__ inline const float* derived_state_rows(const macjd_env_tables& t) {
  const size_t cols = t.env_stride == 0 ? 1 : (size_t)t.n_envs;
  const size_t dbl = (size_t)derived_rows(t.n_jammers, t.n_radars, t.n_targets).total * cols * sizeof(double);
  return reinterpret_cast<const float*>(reinterpret_cast<const char*>(t.derived) + ((dbl + 15) & ~(size_t)15));
}
// lanes per env: the power of two that covers jammers and radars (<= 32)
__host__ __device__ inline int env2_group(int J, int R) {
  int g = 1;
  while (g < J || g < R) g <<= 1;
  return g;
}
inline bool env2_supported(const macjd_env_tables& t) { return t.n_jammers <= 32 && t.n_radars <= 32; }

// ------------------------------------------------------------------------------------------ prepare
__global__ void __launch_bounds__(128) env_prepare_kernel(const macjd_env_tables T, double* out, float* state_rows, int n_cols) {
  grid_dependency_wait();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_cols) return;
  const int J = T.n_jammers, R = T.n_radars, K = T.n_targets;
  const int jbase = 16 * R, tbase = 16 * R + 8 * J;
  const double* col = T.data + (int64_t)e * T.env_stride;
  const int rs = (int)T.row_stride;
  auto tab = [&](int row) -> double { return env_tab(col, rs, row); };
  const DerivedRows D = derived_rows(J, R, K);
  double* blk = out + (int64_t)e * D.total;                            // this env's block
  auto put = [&](int off, double v) { blk[off] = v; };
  const double four_pi3 = (4.0 * 3.141592653589793) * (4.0 * 3.141592653589793) * (4.0 * 3.141592653589793);
  for (int r = 0; r < R; ++r) {
    const int rr = 16 * r;
    // (the expressions of env_step_kernel's radar loop, term for term)
    const double pt = tab(rr + 0), gt = tab(rr + 1), gr = tab(rr + 2);
    const double lam = tab(rr + 3), loss = tab(rr + 4), latm = tab(rr + 5);
    const double pn = tab(rr + 6), Ga = tab(rr + 7), Dd = tab(rr + 8);
    const double rx = tab(rr + 10), ry = tab(rr + 11);
    const double num0 = pt * gt * gr * (lam * lam);
    for (int k = 0; k < K; ++k) {
      const int tr = tbase + 3 * k;
      const double dx = rx - tab(tr + 0), dy = ry - tab(tr + 1);
      const double d = fmax(sqrt(dx * dx + dy * dy), 1e-6);
      const double num = num0 * tab(tr + 2);
      const double d2 = d * d;
      const double den = four_pi3 * (d2 * d2) * loss * latm;
      const double ps = den > 1e-18 ? num / den : 0.0;
      const double sig = Ga * ps;
      const double snr0 = fmax(0.0, pn > 1e-18 ? sig / pn : 0.0);
      put(D.sig + r * K + k, sig);
      put(D.snr0 + r * K + k, snr0);
      put(D.pd0 + r * K + k, albersheim(T, snr0));
    }
    put(D.rad + 4 * r + 0, pn);
    put(D.rad + 4 * r + 1, Dd);
    put(D.rad + 4 * r + 2, fmin(fmax(-tab(rr + 9), T.rd_min), T.rd_max));
    put(D.rad + 4 * r + 3, gr);
    for (int j = 0; j < J; ++j) {
      const int jr = jbase + 8 * j;
      const double dx = tab(jr + 4) - rx, dy = tab(jr + 5) - ry;
      const double dist = sqrt(dx * dx + dy * dy);
      double den = -1.0;                                   // closer than 1e-6: the jammer has no effect on this radar
      if (dist > 1e-6) {
        const double dsq = fmax(1e-9, dist * dist);
        den = dsq * tab(jr + 1) * tab(jr + 2) * fmax(1e-9, tab(jr + 3));
      }
      put(D.den + j * R + r, den);
    }
  }
  for (int j = 0; j < J; ++j) {
    const int jr = jbase + 8 * j;
    put(D.jam + 4 * j + 0, tab(jr + 0));
    put(D.jam + 4 * j + 1, tab(jr + 6));
    put(D.jam + 4 * j + 2, tab(jr + 7));
    put(D.jam + 4 * j + 3, 0.0);
  }
  // float32 state row (environment.py:479-510)
  const int per = 6 + T.n_types, S = R * per + 2 * J;
  float* row = state_rows + (int64_t)e * S;
  for (int r = 0; r < R; ++r) {
    const int rr = 16 * r, o = r * per;
    row[o + 0] = (float)tab(rr + 0);
    row[o + 1] = (float)tab(rr + 12);
    row[o + 2] = (float)tab(rr + 14);
    const int ty = (int)tab(rr + 15);
    for (int c = 0; c < T.n_types; ++c) row[o + 3 + c] = (c == ty) ? 1.0f : 0.0f;
    row[o + 3 + T.n_types] = (float)tab(rr + 13);
    row[o + 4 + T.n_types] = (float)tab(rr + 10);
    row[o + 5 + T.n_types] = (float)tab(rr + 11);
  }
  for (int j = 0; j < J; ++j) {
    row[R * per + 2 * j] = (float)tab(jbase + 8 * j + 4);
    row[R * per + 2 * j + 1] = (float)tab(jbase + 8 * j + 5);
  }
}

inline int env_prepare_launch(const macjd_ctx* ctx, const macjd_env_tables* tab, void* derived) {
  if (!ctx || !tab || !derived || !tab->data) return MACJD_ERR_INVALID_ARG;
  if (tab->n_envs < 0 || tab->n_jammers < 1 || tab->n_radars < 1 || tab->n_targets < 1 || tab->n_types < 1) return MACJD_ERR_INVALID_ARG;
  if (tab->n_radars > 64 || tab->row_stride > 0x7fffffffll || tab->row_stride < 0 || !env2_supported(*tab)) return MACJD_ERR_UNSUPPORTED;
  const int n_cols = tab->env_stride == 0 ? 1 : tab->n_envs;
  if (n_cols == 0) return MACJD_OK;
  macjd_env_tables t = *tab;
  t.derived = reinterpret_cast<const double*>(derived);
  MACJD_LAUNCH(env_prepare_kernel, dim3((n_cols + 127) / 128), dim3(128), 0, (cudaStream_t)ctx->stream, t,
               reinterpret_cast<double*>(derived), const_cast<float*>(derived_state_rows(t)), n_cols);
  return MACJD_OK;
}

// ------------------------------------------------------------------------------------------ step
constexpr int kEnv2Phys = 128;       // physics threads per block (128 / G envs, at most 32)
constexpr int kEnv2Threads = 256;    // + four view warps

struct Env2Args {
  macjd_env_tables tab;
  macjd_env_io io;
  DerivedRows rows;
  const float* state_rows;
  int state_dim, n_actions;
  int physics;               // 0: reset (views only, step_count <- 0)
  int group;                 // lanes per env (env2_group)
  int envs_per_block;        // stand-alone kernel: min(32, 128 / group)
  uint32_t magic_s4, magic_js4;
  int env_begin, env_end;
};

// A launch that runs several timesteps (macjd_rollout_steps) writes the per-step outputs time-major: step t of the
// launch lands t x (size of one step's output) behind step 0's address.  The info-only outputs (pd, detected, ...)
// and the step counters are per-env state and stay where they are.
__device__ __forceinline__ macjd_env_io env2_io_at(const Env2Args& a, int t) {
  macjd_env_io io = a.io;
  if (t == 0) return io;
  const int64_t n = a.tab.n_envs, J = a.tab.n_jammers, S = a.state_dim, A = a.n_actions;
  io.reward += t * n;
  if (io.r_d) io.r_d += t * n;
  if (io.r_p) io.r_p += t * n;
  if (io.r_j) io.r_j += t * n;
  if (io.terminated) io.terminated += t * n;
  if (io.state) io.state += t * n * S;
  if (io.obs) io.obs += t * n * J * S;
  if (io.avail) io.avail += t * n * J * A;
  return io;
}

// One env's step by a group of G lanes (lane g of the group; the groups of a warp run in lockstep: every lane of
// the warp must call this, `live` says whether its env exists).  d: this env's derived block (global or shared
// memory); act_d / act_p: this env's J actions.  Lane 0 of the group writes the per-env outputs.
__device__ __forceinline__ void env2_physics_group(const Env2Args& a, const macjd_env_io& io, int e, bool live, int g, int G,
                                                   const double* d, const int32_t* act_d, const float* act_p) {
  const macjd_env_tables& T = a.tab;
  const DerivedRows& D = a.rows;
  const int n = T.n_envs, J = T.n_jammers, R = T.n_radars, K = T.n_targets, RK = R * K;
  const unsigned full = 0xffffffffu;
  const int lane = (int)(threadIdx.x & 31), base = lane & ~(G - 1);    // first lane of this group inside the warp
  const int step = live ? io.step_count[e] + 1 : 1;                    // environment.py:235
  // Philox yields four uniforms per call: slots 4q .. 4q + 3 share one
  Philox4 px = {0u, 0u, 0u, 0u};
  int px_q = -1;
  auto uniform = [&](int slot) -> float {
    if (io.noise) return io.noise[(int64_t)e * (RK + J) + slot];
    if ((slot >> 2) != px_q) {
      px_q = slot >> 2;
      px = philox4x32_10((uint32_t)px_q, (uint32_t)step, (uint32_t)e, kStreamEnvNoise, (uint32_t)io.seed, (uint32_t)(io.seed >> 32));
    }
    const int q = slot & 3;
    return u01(q == 0 ? px.x : q == 1 ? px.y : q == 2 ? px.z : px.w);
  };
  // ---- jammer g (environment.py:248-302, core/jammer.py:73-98)
  double rp_term = 0.0, val = 0.0;
  int code = 0;                                                        // (target << 2) | {0 nothing, 1 suppression, 2 detected false target}
  if (live && g < J) {
    const int Ti = act_d[g];
    double P = (double)act_p[g];
    P = P < 0.0 ? 0.0 : (P > 1.0 ? 1.0 : P);
    const double gj = d[D.jam + 4 * g], pmin = d[D.jam + 4 * g + 1], pmax = d[D.jam + 4 * g + 2];
    const double range = pmax - pmin;
    const double power = pmin + P * range;
    const double norm = range > 1e-6 ? (power - pmin) / range : 0.0;
    rp_term = T.rp_max + (T.rp_min - T.rp_max) * norm;               // charged even when idle
    if (io.jam_power) io.jam_power[(int64_t)g * n + e] = (float)power;
    if (Ti >= 1 && Ti <= 2 * R && power > 0.0) {
      const int tgt = (Ti + 1) / 2 - 1;
      const double den = d[D.den + g * R + tgt];
      if (den >= 0.0) {                                                // farther than 1e-6 from the radar
        double prj = 0.0;
        if (den > 1e-18) prj = fmax(0.0, (fmax(0.0, power) * gj * d[D.rad + 4 * tgt + 3]) / den);
        if (Ti & 1) {                                                  // suppression
          code = (tgt << 2) | 1; val = prj;
        } else {                                                       // deception: false target (environment.py:408-437)
          const double pn = d[D.rad + 4 * tgt];
          double snr_f = pn > 1e-18 ? (d[D.rad + 4 * tgt + 1] * prj) / pn : 0.0;
          snr_f = fmax(0.0, snr_f);
          const double pd_f = albersheim(T, snr_f);
          if ((double)uniform(RK + g) <= pd_f) { code = (tgt << 2) | 2; val = 1.0 - fmin(pd_f, 0.999999); }
        }
      }
    }
  }
  // ---- the jammers' records reach every lane in jammer order: r_p, and what hits radar g
  double r_p = 0.0, prjs = 0.0, prod = 1.0;
  bool supp = false, hit = false;
  for (int j = 0; j < J; ++j) {
    const double rp_j = __shfl_sync(full, rp_term, base + j);
    const double v_j = __shfl_sync(full, val, base + j);
    const int c_j = __shfl_sync(full, code, base + j);
    r_p += rp_j;
    if ((c_j >> 2) == g) {
      if ((c_j & 3) == 1) { prjs += v_j; supp = true; }
      else if ((c_j & 3) == 2) { prod *= v_j; hit = true; }
    }
  }
  // ---- radar g and its K targets (environment.py:316-349, 359-366, 385-398)
  const bool radar = live && g < R;
  double pn = 0.0, jam = 0.0, den1 = 0.0, rdt = 0.0, red = 0.0;
  if (radar) {
    pn = d[D.rad + 4 * g];
    rdt = d[D.rad + 4 * g + 2];
    if (supp) { jam = d[D.rad + 4 * g + 1] * prjs; den1 = jam + pn; }
  }
  bool tracked = false;
  const bool want_net = io.pd_net != nullptr;
  for (int k = 0; k < K; ++k) {
    double pd = 0.0;
    if (radar) {
      const int slot = g * K + k;
      const double pd0 = d[D.pd0 + slot];
      double sig = 0.0, snr1 = 0.0;
      pd = pd0;
      if (supp || io.jsr_db != nullptr) sig = d[D.sig + slot];
      if (supp) {
        snr1 = den1 > 1e-18 ? sig / den1 : 0.0;
        pd = albersheim(T, snr1);
        red += fmax(0.0, pd0 - pd);                                    // P_d without jamming only matters here (r_j)
      }
      const bool det = (double)uniform(slot) <= pd;
      tracked |= det;
      const int64_t o = (int64_t)slot * n + e;
      if (io.pd) io.pd[o] = (float)pd;
      if (io.detected) io.detected[o] = det ? 1 : 0;
      if (io.snr0 || io.snr1) {
        const float s0 = (float)d[D.snr0 + slot];
        if (io.snr0) io.snr0[o] = s0;
        if (io.snr1) io.snr1[o] = supp ? (float)fmax(0.0, snr1) : s0;
      }
      if (io.jsr_db) io.jsr_db[o] = 10.0f * log10f((float)(jam / sig));   // float32 output of an extension: float log
    }
    if (want_net) {                                                    // prod_r (1 - pd[r][k]) in radar order
      double pn_k = 1.0;
      const double mine = 1.0 - pd;
      for (int r = 0; r < R; ++r) pn_k *= __shfl_sync(full, mine, base + r);
      if (live && g == 0) io.pd_net[(int64_t)k * n + e] = (float)(1.0 - pn_k);
    }
  }
  if (radar && io.tracking) io.tracking[(int64_t)g * n + e] = tracked ? 1 : 0;
  // ---- the per-radar terms, summed in radar order (a term the sequential code skips is an exact + 0.0 here)
  const double t_d = (radar && tracked) ? rdt : 0.0, t_s = (radar && supp) ? red : 0.0, t_h = (radar && hit) ? 1.0 - prod : 0.0;
  double r_d = 0.0, r_j_supp = 0.0, r_j_dec = 0.0;
  for (int r = 0; r < R; ++r) {
    r_d += __shfl_sync(full, t_d, base + r);
    r_j_supp += __shfl_sync(full, t_s, base + r);
    r_j_dec += __shfl_sync(full, t_h, base + r);
  }
  if (live && g == 0) {
    const double r_j = r_j_supp + r_j_dec;
    const double reward = r_d + r_p + r_j;                             // environment.py:457
    const bool term = step >= T.episode_limit;                         // environment.py:460
    io.reward[e] = (float)reward;
    if (io.reward64) io.reward64[e] = reward;
    if (io.r_d) io.r_d[e] = (float)r_d;
    if (io.r_p) io.r_p[e] = (float)r_p;
    if (io.r_j) io.r_j[e] = (float)r_j;
    if (io.terminated) io.terminated[e] = term ? 1 : 0;
    io.step_count[e] = (term && io.auto_reset) ? 0 : step;
  }
}

// The static views (environment.py:479-551) of envs [e0, e0 + valid): thread vt of VT cooperating threads.
__device__ __forceinline__ void env2_views(const Env2Args& a, const macjd_env_io& io, int e0, int valid, int vt, int VT) {
  const macjd_env_tables& T = a.tab;
  const int J = T.n_jammers, S = a.state_dim, A = a.n_actions;
  const bool shared_scn = T.env_stride == 0;
  if (valid <= 0) return;
  const float* src_rows = a.state_rows + (shared_scn ? 0 : (int64_t)e0 * S);
  if ((S & 3) == 0) {
    const int S4 = S >> 2, JS4 = J * S4;
    const float4* src = reinterpret_cast<const float4*>(src_rows);
    if (io.state) {
      float4* dst = reinterpret_cast<float4*>(io.state + (int64_t)e0 * S);
      for (int v = vt; v < valid * S4; v += VT) {
        const int el = a.magic_s4 ? (int)__umulhi((uint32_t)v, a.magic_s4) : v / S4;
        dst[v] = __ldg(src + (shared_scn ? v - el * S4 : v));
      }
    }
    if (io.obs) {
      float4* dst = reinterpret_cast<float4*>(io.obs + (int64_t)e0 * J * S);
      const int total = valid * JS4;
      // four independent 16-byte copies in flight per thread
      for (int v0 = vt; v0 < total; v0 += 4 * VT) {
        float4 q[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int v = v0 + i * VT;
          if (v < total) {
            const int el = a.magic_js4 ? (int)__umulhi((uint32_t)v, a.magic_js4) : v / JS4;
            const int w = v - el * JS4;
            const int s4 = w - (a.magic_s4 ? (int)__umulhi((uint32_t)w, a.magic_s4) : w / S4) * S4;
            q[i] = __ldg(src + (shared_scn ? 0 : el * S4) + s4);
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int v = v0 + i * VT;
          if (v < total) dst[v] = q[i];
        }
      }
    }
  } else {
    if (io.state) {
      float* dst = io.state + (int64_t)e0 * S;
      for (int v = vt; v < valid * S; v += VT) { const int el = v / S; dst[v] = __ldg(src_rows + (shared_scn ? v - el * S : v)); }
    }
    if (io.obs) {
      const int JS = J * S;
      float* dst = io.obs + (int64_t)e0 * JS;
      for (int v = vt; v < valid * JS; v += VT) { const int el = v / JS; dst[v] = __ldg(src_rows + (shared_scn ? 0 : el * S) + (v % S)); }
    }
  }
  if (io.avail) {                      // all actions always available (environment.py:539-551)
    uint8_t* base = io.avail + (int64_t)e0 * J * A;
    const int total = valid * J * A;
    // 16-byte stores over the aligned middle, bytes at the ragged ends
    const int head = min(total, (int)((16 - (reinterpret_cast<uintptr_t>(base) & 15)) & 15));
    for (int v = vt; v < head; v += VT) base[v] = 1;
    const int n16 = (total - head) >> 4;
    uint4* mid = reinterpret_cast<uint4*>(base + head);
    const uint4 ones = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
    for (int v = vt; v < n16; v += VT) mid[v] = ones;
    for (int v = head + (n16 << 4) + vt; v < total; v += VT) base[v] = 1;
  }
}

// kUniform (large batches): every thread of the block is a physics lane AND a view thread -- 256 / G envs per block,
// each warp first copies its share of the block's static views (independent loads and stores, fire and forget) and
// then runs its envs' physics (a dependent FP64 chain).  With the split roles a block keeps 2-4 physics warps alive
// for the length of that chain and the SM holds 128 envs in flight: at 1 M envs the kernel was bound by
// latency x occupancy (ncu r2s: warps active 30 %, long_scoreboard + wait), not by HBM.  The split form stays for
// small batches, where the step is one chain long and the views should not sit in front of it.
template <bool kUniform>
__global__ void __launch_bounds__(kEnv2Threads) env_step2_kernel(const Env2Args a) {
  const macjd_env_tables& T = a.tab;
  const macjd_env_io& io = a.io;
  const int J = T.n_jammers;
  const int tid = (int)threadIdx.x;
  const int G = a.group, eb = a.envs_per_block;
  const int e0 = a.env_begin + blockIdx.x * eb;
  const int valid = min(eb, a.env_end - e0);
  const int pf_end = (io.snr0 || io.snr1) ? a.rows.total : a.rows.snr0;   // what a step reads of an env's block
  if (kUniform) {
    const int slot = tid / G, g = tid - slot * G;                          // (eb * G == kEnv2Threads)
    const bool live = slot < valid;
    const int e = e0 + (live ? slot : 0);
    const double* d = T.derived + (T.env_stride == 0 ? 0 : (int64_t)e * a.rows.total);
    if (a.physics && live)
      for (int off = g * 16; off < pf_end; off += G * 16) prefetch_l2(d + off);
    grid_dependency_wait();
    env2_views(a, io, e0, valid, tid, kEnv2Threads);
    if (!a.physics) {
      if (live && g == 0) io.step_count[e] = 0;                            // environment.py:203
      return;
    }
    env2_physics_group(a, io, e, live, g, G, d, io.act_d + (int64_t)e * J, io.act_p + (int64_t)e * J);
    return;
  }
  if (tid < kEnv2Phys) {
    // =========================================================================== physics: G lanes per env
    if (tid >= eb * G) return;                                         // (whole warps: eb * G is a multiple of 32)
    const int slot = tid / G, g = tid - slot * G;
    const bool live = slot < valid;
    const int e = e0 + (live ? slot : 0);
    const double* d = T.derived + (T.env_stride == 0 ? 0 : (int64_t)e * a.rows.total);
    // pull this env's block towards the SM before waiting for the actions (the table is read-only)
    if (a.physics && live)
      for (int off = g * 16; off < pf_end; off += G * 16) prefetch_l2(d + off);
    // from here on the kernel reads the actions and writes outputs: wait for the preceding kernel of the stream
    grid_dependency_wait();
    if (!a.physics) {
      if (live && g == 0) io.step_count[e] = 0;                        // environment.py:203
      return;
    }
    env2_physics_group(a, io, e, live, g, G, d, io.act_d + (int64_t)e * J, io.act_p + (int64_t)e * J);
    return;
  }
  // ============================================================================= views: warps 4-7
  // (nothing here depends on the actions, but the destinations may be read by the preceding kernel)
  grid_dependency_wait();
  env2_views(a, io, e0, valid, tid - kEnv2Phys, kEnv2Threads - kEnv2Phys);
}

// kernel arguments of one step (also used by the agent kernel when it runs the step itself)
inline Env2Args env2_args(const macjd_env_tables* tab, const macjd_env_io* io, int physics) {
  const int n_step = io->env_count > 0 ? io->env_count : tab->n_envs - io->env_begin;
  Env2Args a;
  a.tab = *tab;
  a.io = *io;
  a.rows = derived_rows(tab->n_jammers, tab->n_radars, tab->n_targets);
  a.state_rows = derived_state_rows(*tab);
  a.state_dim = tab->n_radars * (6 + tab->n_types) + 2 * tab->n_jammers;
  a.n_actions = 2 * tab->n_radars + 1;
  a.physics = physics;
  a.group = env2_group(tab->n_jammers, tab->n_radars);
  a.envs_per_block = kEnv2Phys / a.group < 32 ? kEnv2Phys / a.group : 32;
  a.env_begin = io->env_begin;
  a.env_end = io->env_begin + n_step;
  {
    const uint64_t s4 = (uint64_t)a.state_dim / 4, js4 = (uint64_t)tab->n_jammers * s4;
    const bool ok = (a.state_dim % 4 == 0) && s4 >= 2 && 128ull * js4 * js4 < 0x100000000ull;   // (<= 128 envs per copying block)
    a.magic_s4 = ok ? (uint32_t)((0x100000000ull + s4 - 1) / s4) : 0;
    a.magic_js4 = ok ? (uint32_t)((0x100000000ull + js4 - 1) / js4) : 0;
  }
  return a;
}

// batch size from which the stand-alone kernel runs in its uniform form (MACJD_ENV_UNIFORM_MIN overrides: tests, experiments)
inline int env2_uniform_min() {
  const char* e = getenv("MACJD_ENV_UNIFORM_MIN");          // read per call (tens of ns): tests switch it
  return e ? atoi(e) : 16384;
}

inline int env2_launch(const macjd_ctx* ctx, const macjd_env_tables* tab, const macjd_env_io* io, int physics) {
  if (!env2_supported(*tab)) return MACJD_ERR_UNSUPPORTED;
  const int n_step = io->env_count > 0 ? io->env_count : tab->n_envs - io->env_begin;
  if (n_step == 0) return MACJD_OK;
  Env2Args a = env2_args(tab, io, physics);
  if (n_step >= env2_uniform_min()) {
    a.envs_per_block = kEnv2Threads / a.group;
    MACJD_LAUNCH(env_step2_kernel<true>, (n_step + a.envs_per_block - 1) / a.envs_per_block, kEnv2Threads, 0, (cudaStream_t)ctx->stream, a);
    return MACJD_OK;
  }
  const int grid = (n_step + a.envs_per_block - 1) / a.envs_per_block;
  MACJD_LAUNCH(env_step2_kernel<false>, grid, kEnv2Threads, 0, (cudaStream_t)ctx->stream, a);
  return MACJD_OK;
}

// environment.py:221-477 / :208-219 for the whole batch: on the derived tables when the caller prepared them
inline int env_launch(const macjd_ctx* ctx, const macjd_env_tables* tab, const macjd_env_io* io, int physics) {
  const int st = env_check_args(ctx, tab, io, physics);
  if (st != MACJD_OK) return st;
  return tab->derived ? env2_launch(ctx, tab, io, physics) : env_launch_raw(ctx, tab, io, physics);
}

}  // namespace macjd
