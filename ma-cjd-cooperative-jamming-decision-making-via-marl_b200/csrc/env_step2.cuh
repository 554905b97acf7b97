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
// The step itself then costs one division per acting jammer, one Albersheim evaluation per deception attempt
// and per target of a SUPPRESSED radar, and the Bernoulli draws: ~20 exponentials per env-step instead of 84
// at 8 jammers x 16 radars x 4 targets, and a short dependent chain at the default scenario.  Sums and products
// keep the reference's order (jammers in jammer order into each radar, radars in radar order), so the
// outputs are those of env_step_kernel (env_step.cuh, kept for callers without derived tables).
//
// A block is 128 threads = 32 envs: warp 0 runs the physics (thread = env; outputs are [field][env], coalesced),
// warps 1-3 copy the static views of the same 32 envs (state, obs = the state once per jammer, all-ones
// availability: 6.3 KB per env-step at 8 x 16 x 4, the bulk of the kernel's traffic) with 16-byte stores,
// concurrently.  Launched behind the agent kernel (programmatic dependent launch) the physics warp has the env's
// derived rows in shared memory before it waits for the actions.
#pragma once
#include "env_step.cuh"

namespace macjd {

// row indices of the derived table (doubles, [row][env] like the raw tables)
struct DerivedRows {
  int sig, pd0, snr0;      // [R*K] each
  int pn, dd, rdt, gr;     // [R] each
  int den;                 // [J*R]
  int gj, pmin, pmax;      // [J] each
  int total;
};
__host__ __device__ inline DerivedRows derived_rows(int J, int R, int K) {
  DerivedRows d;
  int o = 0;
  d.sig = o; o += R * K; d.pd0 = o; o += R * K; d.snr0 = o; o += R * K;
  d.pn = o; o += R; d.dd = o; o += R; d.rdt = o; o += R; d.gr = o; o += R;
  d.den = o; o += J * R;
  d.gj = o; o += J; d.pmin = o; o += J; d.pmax = o; o += J;
  d.total = o;
  return d;
}
inline size_t env_derived_bytes(const macjd_env_tables& t) {
  const size_t cols = t.env_stride == 0 ? 1 : (size_t)t.n_envs;
  const size_t S = (size_t)t.n_radars * (6 + t.n_types) + 2 * (size_t)t.n_jammers;
  const size_t dbl = (size_t)derived_rows(t.n_jammers, t.n_radars, t.n_targets).total * cols * sizeof(double);
  return ((dbl + 15) & ~(size_t)15) + ((cols * S * sizeof(float) + 15) & ~(size_t)15);
}
// the float32 state rows sit behind the double rows
__host__ __device__ inline const float* derived_state_rows(const macjd_env_tables& t) {
  const size_t cols = t.env_stride == 0 ? 1 : (size_t)t.n_envs;
  const size_t dbl = (size_t)derived_rows(t.n_jammers, t.n_radars, t.n_targets).total * cols * sizeof(double);
  return reinterpret_cast<const float*>(reinterpret_cast<const char*>(t.derived) + ((dbl + 15) & ~(size_t)15));
}

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
  const int64_t ors = T.env_stride == 0 ? 1 : (int64_t)T.n_envs;       // same [row][env] convention as the raw table
  auto put = [&](int row, double v) { out[(int64_t)row * ors + (T.env_stride == 0 ? 0 : e)] = v; };
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
    put(D.pn + r, pn);
    put(D.dd + r, Dd);
    put(D.gr + r, gr);
    put(D.rdt + r, fmin(fmax(-tab(rr + 9), T.rd_min), T.rd_max));
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
    put(D.gj + j, tab(jr + 0));
    put(D.pmin + j, tab(jr + 6));
    put(D.pmax + j, tab(jr + 7));
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
  if (tab->n_radars > 64 || tab->row_stride > 0x7fffffffll || tab->row_stride < 0) return MACJD_ERR_UNSUPPORTED;
  const int n_cols = tab->env_stride == 0 ? 1 : tab->n_envs;
  if (n_cols == 0) return MACJD_OK;
  macjd_env_tables t = *tab;
  t.derived = reinterpret_cast<const double*>(derived);
  MACJD_LAUNCH(env_prepare_kernel, dim3((n_cols + 127) / 128), dim3(128), 0, (cudaStream_t)ctx->stream, t,
               reinterpret_cast<double*>(derived), const_cast<float*>(derived_state_rows(t)), n_cols);
  return MACJD_OK;
}

// ------------------------------------------------------------------------------------------ step
constexpr int kEnv2Envs = 32;        // envs per block (the physics warp)
constexpr int kEnv2Threads = 128;    // + three view warps

struct Env2Args {
  macjd_env_tables tab;
  macjd_env_io io;
  DerivedRows rows;
  const float* state_rows;
  int state_dim, n_actions;
  int physics;               // 0: reset (views only, step_count <- 0)
  int stage_rows;            // physics warp copies its envs' derived rows to shared memory before the dependency wait
  uint32_t magic_s4, magic_js4;
  int env_begin, env_end;
};

// One env's step on the derived tables.  `dv(row)` reads this env's derived value; act_d / act_p point at this env's
// J actions (global memory, or shared memory when the agent kernel runs the step itself); scratch is per-thread
// [slot * sstride + sidx]: rec_val [J], rec_code [J] (ints), pnet [K] (only touched when pd_net is wanted).
template <typename DV>
__device__ __forceinline__ void env2_physics(const Env2Args& a, const macjd_env_io& io, int e, DV dv, const int32_t* act_d,
                                             const float* act_p, double* rec_val, int* rec_code, double* pnet, int sstride, int sidx) {
  const macjd_env_tables& T = a.tab;
  const DerivedRows& D = a.rows;
  const int n = T.n_envs, J = T.n_jammers, R = T.n_radars, K = T.n_targets, RK = R * K;
  const int step = io.step_count[e] + 1;                             // environment.py:235
  double r_p = 0.0, r_d = 0.0, r_j_supp = 0.0, r_j_dec = 0.0;
  uint64_t supp_mask = 0, hit_mask = 0;
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
  // ---- jammers (environment.py:248-302, core/jammer.py:73-98), in jammer order
  for (int j = 0; j < J; ++j) {
    const int Ti = act_d[j];
    double P = (double)act_p[j];
    P = P < 0.0 ? 0.0 : (P > 1.0 ? 1.0 : P);
    const double pmin = dv(D.pmin + j), pmax = dv(D.pmax + j);
    const double range = pmax - pmin;
    const double power = pmin + P * range;
    const double norm = range > 1e-6 ? (power - pmin) / range : 0.0;
    r_p += T.rp_max + (T.rp_min - T.rp_max) * norm;                // charged even when idle
    if (io.jam_power) io.jam_power[(int64_t)j * n + e] = (float)power;
    int code = 0;
    double val = 0.0;
    if (Ti >= 1 && Ti <= 2 * R && power > 0.0) {
      const int tgt = (Ti + 1) / 2 - 1;
      const double den = dv(D.den + j * R + tgt);
      if (den >= 0.0) {                                              // farther than 1e-6 from the radar
        double prj = 0.0;
        if (den > 1e-18) prj = fmax(0.0, (fmax(0.0, power) * dv(D.gj + j) * dv(D.gr + tgt)) / den);
        if (Ti & 1) {                                                // suppression
          code = (tgt << 2) | 1; val = prj;
          supp_mask |= 1ull << tgt;
        } else {                                                     // deception: false target (environment.py:408-437)
          const double pn = dv(D.pn + tgt);
          double snr_f = pn > 1e-18 ? (dv(D.dd + tgt) * prj) / pn : 0.0;
          snr_f = fmax(0.0, snr_f);
          const double pd_f = albersheim(T, snr_f);
          if ((double)uniform(RK + j) <= pd_f) {
            code = (tgt << 2) | 2; val = 1.0 - fmin(pd_f, 0.999999);
            hit_mask |= 1ull << tgt;
          }
        }
      }
    }
    rec_val[j * sstride + sidx] = val;
    rec_code[j * sstride + sidx] = code;
  }
  const bool want_net = io.pd_net != nullptr;
  if (want_net)
    for (int k = 0; k < K; ++k) pnet[k * sstride + sidx] = 1.0;
  // ---- radars x targets (environment.py:316-349, 359-366, 385-398), in radar order
  for (int r = 0; r < R; ++r) {
    const bool supp = (supp_mask >> r) & 1ull, hit = (hit_mask >> r) & 1ull;
    double prjs = 0.0, prod = 1.0;
    if (supp || hit)
      for (int j = 0; j < J; ++j) {                                  // this radar's jammers, in jammer order
        const int code = rec_code[j * sstride + sidx];
        if ((code >> 2) == r) {
          if ((code & 3) == 1) prjs += rec_val[j * sstride + sidx];
          else if ((code & 3) == 2) prod *= rec_val[j * sstride + sidx];
        }
      }
    double jam = 0.0, den1 = 0.0;
    if (supp) { jam = dv(D.dd + r) * prjs; den1 = jam + dv(D.pn + r); }
    bool tracked = false;
    double red = 0.0;
    for (int k = 0; k < K; ++k) {
      const int slot = r * K + k;
      const double pd0 = dv(D.pd0 + slot);
      double pd = pd0, sig = 0.0, snr1 = 0.0;
      const bool need_sig = supp || io.jsr_db != nullptr;
      if (need_sig) sig = dv(D.sig + slot);
      if (supp) {
        snr1 = den1 > 1e-18 ? sig / den1 : 0.0;
        pd = albersheim(T, snr1);
        red += fmax(0.0, pd0 - pd);                                  // P_d without jamming only matters here (r_j)
      }
      const bool det = (double)uniform(slot) <= pd;
      tracked |= det;
      if (want_net) pnet[k * sstride + sidx] *= (1.0 - pd);
      const int64_t o = (int64_t)slot * n + e;
      if (io.pd) io.pd[o] = (float)pd;
      if (io.detected) io.detected[o] = det ? 1 : 0;
      if (io.snr0 || io.snr1) {
        const float s0 = (float)dv(D.snr0 + slot);
        if (io.snr0) io.snr0[o] = s0;
        if (io.snr1) io.snr1[o] = supp ? (float)fmax(0.0, snr1) : s0;
      }
      if (io.jsr_db) io.jsr_db[o] = 10.0f * log10f((float)(jam / sig));   // float32 output of an extension: float log
    }
    if (io.tracking) io.tracking[(int64_t)r * n + e] = tracked ? 1 : 0;
    if (tracked) r_d += dv(D.rdt + r);
    if (supp) r_j_supp += red;
    if (hit) r_j_dec += 1.0 - prod;
  }
  if (want_net)
    for (int k = 0; k < K; ++k) io.pd_net[(int64_t)k * n + e] = (float)(1.0 - pnet[k * sstride + sidx]);
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

template <bool kStaged>
__global__ void __launch_bounds__(kEnv2Threads) env_step2_kernel(const Env2Args a) {
  const macjd_env_tables& T = a.tab;
  const macjd_env_io& io = a.io;
  const DerivedRows& D = a.rows;
  const int n = T.n_envs, J = T.n_jammers, K = T.n_targets;
  const int tid = (int)threadIdx.x;
  const int e0 = a.env_begin + blockIdx.x * kEnv2Envs;
  const int valid = min(kEnv2Envs, a.env_end - e0);
  const bool shared_scn = T.env_stride == 0;
  MACJD_DYNAMIC_SMEM(double, smem);

  if (tid < kEnv2Envs) {
    // =========================================================================== physics: thread = env
    const int e = e0 + tid;
    const bool live = tid < valid;
    const int64_t drs = shared_scn ? 1 : (int64_t)n;                   // derived row stride
    const double* dcol = T.derived + (shared_scn ? 0 : (live ? e : e0));
    // per-thread scratch [slot][32]: jammer records (value, code), networked-Pd products
    double* rec_val = smem;                                            // [J][32]
    int* rec_code = reinterpret_cast<int*>(rec_val + (size_t)J * kEnv2Envs);   // [J][32]
    double* pnet = rec_val + (size_t)J * kEnv2Envs + ((size_t)J * kEnv2Envs + 1) / 2;   // [K][32]
    double* srows = pnet + (size_t)K * kEnv2Envs;                      // [rows][32] when kStaged
    if (kStaged && a.physics && live) {
#pragma unroll 1
      for (int row0 = 0; row0 < D.total; row0 += 16) {
        double v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = (row0 + i < D.total) ? __ldg(dcol + (int64_t)(row0 + i) * drs) : 0.0;
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (row0 + i < D.total) srows[(row0 + i) * kEnv2Envs + tid] = v[i];
      }
    }
    auto dv = [&](int row) -> double { return kStaged ? srows[row * kEnv2Envs + tid] : __ldg(dcol + (int64_t)row * drs); };
    // from here on the kernel reads the actions and writes outputs: wait for the preceding kernel of the stream
    grid_dependency_wait();
    if (live && !a.physics) io.step_count[e] = 0;                      // environment.py:203
    if (live && a.physics)
      env2_physics(a, io, e, dv, io.act_d + (int64_t)e * J, io.act_p + (int64_t)e * J, rec_val, rec_code, pnet, kEnv2Envs, tid);
    return;
  }

  // ============================================================================= views: warps 1-3, 32 envs
  // (nothing here depends on the actions, but the destinations may be read by the preceding kernel: wait before
  // the first store)
  grid_dependency_wait();
  env2_views(a, io, e0, valid, tid - kEnv2Envs, kEnv2Threads - kEnv2Envs);
}

inline size_t env2_smem_bytes(int J, int K, int staged_rows) {
  return ((size_t)J * kEnv2Envs + ((size_t)J * kEnv2Envs + 1) / 2 + (size_t)K * kEnv2Envs + (size_t)staged_rows * kEnv2Envs) * sizeof(double);
}

// kernel arguments of one step (also used by the agent kernel when it runs the step itself)
inline Env2Args env2_args(const macjd_env_tables* tab, const macjd_env_io* io, int physics);

inline int env2_launch(const macjd_ctx* ctx, const macjd_env_tables* tab, const macjd_env_io* io, int physics) {
  const int n_step = io->env_count > 0 ? io->env_count : tab->n_envs - io->env_begin;
  if (n_step == 0) return MACJD_OK;
  Env2Args a = env2_args(tab, io, physics);
  // Small scenarios: the physics warp parks its envs' derived rows in shared memory ahead of the dependency wait
  // (behind the agent kernel that wait is long and the copy free; every later lookup is a shared-memory read on the
  // dependent chain).  Large scenarios read the few rows a step needs straight from L2.
  a.stage_rows = (physics && (io->flags & MACJD_ENV_FOLLOWS_AGENT) && a.rows.total <= 96 && tab->env_stride != 0) ? 1 : 0;
  const size_t smem = env2_smem_bytes(tab->n_jammers, tab->n_targets, a.stage_rows ? a.rows.total : 0);
  if (smem > 48 * 1024) return MACJD_ERR_UNSUPPORTED;
  const int grid = (n_step + kEnv2Envs - 1) / kEnv2Envs;
  if (a.stage_rows) MACJD_LAUNCH(env_step2_kernel<true>, grid, kEnv2Threads, smem, (cudaStream_t)ctx->stream, a);
  else MACJD_LAUNCH(env_step2_kernel<false>, grid, kEnv2Threads, smem, (cudaStream_t)ctx->stream, a);
  return MACJD_OK;
}

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
  a.env_begin = io->env_begin;
  a.env_end = io->env_begin + n_step;
  {
    const uint64_t s4 = (uint64_t)a.state_dim / 4, js4 = (uint64_t)tab->n_jammers * s4;
    const bool ok = (a.state_dim % 4 == 0) && s4 >= 2 && 64ull * js4 * js4 < 0x100000000ull;   // (<= 64 envs per copying block)
    a.magic_s4 = ok ? (uint32_t)((0x100000000ull + s4 - 1) / s4) : 0;
    a.magic_js4 = ok ? (uint32_t)((0x100000000ull + js4 - 1) / js4) : 0;
  }
  a.stage_rows = 0;
  return a;
}

// environment.py:221-477 / :208-219 for the whole batch: on the derived tables when the caller prepared them
inline int env_launch(const macjd_ctx* ctx, const macjd_env_tables* tab, const macjd_env_io* io, int physics) {
  const int st = env_check_args(ctx, tab, io, physics);
  if (st != MACJD_OK) return st;
  return tab->derived ? env2_launch(ctx, tab, io, physics) : env_launch_raw(ctx, tab, io, physics);
}

}  // namespace macjd
