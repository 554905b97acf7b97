// Fused batched environment step (K1 of SURVEY.md): one thread per episode.
//
// For every env the kernel evaluates what simulation/environment.py:221-477 does for its
// single instance: action decode, jammer power bookkeeping and r_p, Friis jamming power
// per jammer->radar link (core/jammer.py:56-98), radar-equation echo power per
// radar-target pair (core/radar.py:35-60), SNR with / without suppression jamming,
// Albersheim Pd (core/radar.py:67-82), Monte-Carlo detection and false-target draws,
// r_d / r_j, termination -- and rewrites the static views (state, obs, avail-action
// mask; environment.py:479-551).  Extensions: K > 1 targets, JSR in dB,
// networked Pd.
//
// Numerics: float64, the reference's own number type (it decides `u <= pd` and the
// catastrophically cancelling pd0 - pd1 of r_j), outputs rounded to float32 on store.
// B200 has full-rate-class FP64 (64 lanes/SM); the kernel stays HBM-bound.
//
// Memory: scenario tables are struct-of-arrays [row][env] so a warp's table loads are
// contiguous; per-env scalars out are [field][env]; the row-major views the agent
// kernel consumes (state [env][S], obs [env][J][S], avail [env][J][A]) are staged in
// shared memory and written back with coalesced 16-byte stores.
#pragma once
#include "macjd_common.cuh"

namespace macjd {

struct EnvKernelArgs {
  macjd_env_tables tab;
  macjd_env_io io;
  int state_dim;      // S = R*(6+types) + 2J
  int n_actions;      // A = 2R+1
  int stage_ld;       // padded row length of the smem state stage (odd -> no bank conflicts)
  int physics;        // 0: reset (views only, step_count <- 0), 1: full step
  int prefetch;       // issue L2 prefetches for the env's table column first (small batches)
  int stage_tab;      // small batches, launched right behind the kernel that produces the actions (io.flags):
                      // copy the scenario column to shared memory while waiting for it
  int split_views;    // small batches: a second set of warps writes the static views while the first runs
                      // the physics (the block is launched with 2 x the env count; one dependent chain
                      // instead of two in sequence)
  uint32_t magic_s4;  // ceil(2^32 / (S/4)), ceil(2^32 / (J*S/4)): exact v / d for the view write-back loops
  uint32_t magic_js4;
  int env_begin, env_end;   // this launch steps envs [env_begin, env_end) (io.env_begin / env_count; default all)
};

// element `row` of one env's table column (col = data + env * env_stride; row stride fits 32 bits)
__device__ __forceinline__ double env_tab(const double* col, int row_stride, int row) {
  return __ldg(col + (int64_t)row * row_stride);
}

__device__ __forceinline__ double albersheim(const macjd_env_tables& t, double snr) {
  // core/radar.py:67-82 (prfa, m folded into alb_a / alb_zoff / alb_den on the host)
  const double s = snr > 0.0 ? snr : 0.0;
  const double z = s + t.alb_zoff;
  const double b = (10.0 * z - t.alb_a) / t.alb_den;
  if (b > 700.0) return 1.0;
  if (b < -700.0) return 0.0;
  return 1.0 / (1.0 + exp(-b));
}

// Optional %globaltimer stamps of block 0 / thread 0 (-DMACJD_TC_PROFILE; tools/env_phase_profile.py)
#if defined(MACJD_TC_PROFILE) && !defined(MACJD_TEST_HOST_EMULATION)
__device__ unsigned long long g_env_prof[16];
#define ENV_STAMP(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_env_prof[k] = t_; } } while (0)
#else
#define ENV_STAMP(k) do { } while (0)
#endif

// kMode 0: one thread per env does everything.  Small batches are one long dependent FP64 chain per
// thread (the time for 32 envs equals the time for 4 096), so they split it over concurrent warp sets:
// kMode 1: the block has 2 BS threads -- set 0 runs the physics, set 1 writes the static views of the
//          same envs;
// kMode 2: 3 BS threads -- sets 0 and 1 both run the (short) jammer loop, then take the even / odd
//          radars; set 0 adds the per-radar terms up in radar order (so every sum has the sequential
//          kernel's order and value); set 2 writes the views.
// kStaged (kMode 1 / 2 only): the launch follows the agent step; the scenario column is copied to shared
// memory ahead of the dependency wait (compile-time, so that the physics carries no second lookup path).
template <int kMode, bool kStaged = false>
__global__ void __launch_bounds__(kMode == 2 ? 384 : kMode == 1 ? 256 : 128, kMode == 0 ? 5 : 2) env_step_kernel(const EnvKernelArgs a) {
  const macjd_env_tables& T = a.tab;
  const macjd_env_io& io = a.io;
  const int n = T.n_envs, J = T.n_jammers, R = T.n_radars, K = T.n_targets;
  const int RK = R * K, S = a.state_dim, A = a.n_actions;
  constexpr int NW = kMode == 2 ? 2 : 1;                     // physics workers per env
  constexpr int kSets = kMode == 0 ? 1 : NW + 1;
  const int BS = (int)blockDim.x / kSets;                    // envs per block
  const int set_id = kMode == 0 ? 0 : (int)threadIdx.x / BS;
  const bool do_phys = kMode == 0 || set_id < NW;
  const bool do_views = kMode == 0 || set_id == NW;
  const int wk = (kMode == 2 && do_phys) ? set_id : 0;       // which physics worker
  const int tid = (int)threadIdx.x - set_id * BS;
  const int e0 = a.env_begin + blockIdx.x * BS;
  const int e = e0 + tid;
  const bool live = e < a.env_end;
  const int rbase = 0, jbase = 16 * R, tbase = 16 * R + 8 * J;

  MACJD_DYNAMIC_SMEM(double, smem);
  // per-thread scratch, [slot][thread] so that a warp touches consecutive banks
  double* prjs = smem + (size_t)wk * 2 * R * BS;     // [R][BS] accumulated suppression power per radar (per worker)
  double* prod = prjs + (size_t)R * BS;              // [R][BS] prod(1 - pd_f) over detected false targets
  double* pnet = smem + (size_t)NW * 2 * R * BS;     // [K][BS] prod_r (1 - pd[r][k])
  float* stage = reinterpret_cast<float*>(smem + (size_t)(NW * 2 * R + K) * BS);  // [BS][stage_ld]
  // kMode 2: per-radar terms handed to worker 0: [3][R][BS] (r_d, r_j suppression, r_j deception), [RK][BS] pd
  double* rad = smem + (size_t)(NW * 2 * R + K) * BS + ((size_t)BS * a.stage_ld + 1) / 2;
  double* pdv = rad + (size_t)3 * R * BS;

  // The physics below walks the scenario tables with data-dependent, serial lookups.  Request
  // this env's whole table column (and its action / noise rows) up front so that the chain runs
  // against L2 instead of DRAM; a warp's requests cover contiguous 256-byte row segments.
  const double* col = T.data + (int64_t)(live ? e : 0) * T.env_stride;   // this env's table column
  const int rs = (int)T.row_stride;
  if (live && a.prefetch && do_phys) {
    const int n_rows = 16 * R + 8 * J + 3 * K;
    for (int row = 0; row < n_rows; ++row) prefetch_l2(col + (int64_t)row * rs);
    if (a.physics) {
      prefetch_l2(io.act_d + (int64_t)e * J);
      prefetch_l2(io.act_p + (int64_t)e * J);
      if (io.noise) prefetch_l2(io.noise + (int64_t)e * (RK + J));
    }
  }

  // Small-batch kernels: this env's whole scenario column goes to shared memory BEFORE the dependency wait
  // (the tables are read-only, and while the agent kernel runs this costs nothing); after the wait every
  // table lookup of the physics is a shared-memory read instead of an L2 round trip on the dependent chain
  // (jammer loop 4.7 -> us, tools/env_phase_profile.py).  16 loads in flight per thread; the workers of an
  // env take alternate rows and meet at a named barrier.
  const int n_tab_rows = 16 * R + 8 * J + 3 * K;
  double* tabs = pdv + (size_t)RK * BS;                  // [n_tab_rows][BS], kMode 1 / 2 only
  constexpr bool staged = kMode != 0 && kStaged;
  if (staged && a.physics && do_phys) {
    if (live) {
#pragma unroll 1
      for (int row0 = wk * 16; row0 < n_tab_rows; row0 += 16 * NW) {
        double v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = (row0 + i < n_tab_rows) ? env_tab(col, rs, row0 + i) : 0.0;
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (row0 + i < n_tab_rows) tabs[(row0 + i) * BS + tid] = v[i];
      }
    }
#ifndef MACJD_TEST_HOST_EMULATION
    asm volatile("bar.sync 2, %0;\n" ::"r"(NW * BS) : "memory");
#endif
  }
  auto tab = [&](int row) -> double { return staged ? tabs[row * BS + tid] : env_tab(col, rs, row); };

  const bool phys = live && a.physics && do_phys;
  int step = 0;
  double r_p = 0.0, r_d = 0.0, r_j_supp = 0.0, r_j_dec = 0.0;
  // Everything above reads the scenario tables only.  From here on the kernel reads the actions and
  // writes outputs: wait for the preceding kernel of the stream (normally the agent step that chose
  // the actions) -- a no-op unless this launch was allowed to start early (env_launch).
  ENV_STAMP(0);
  if (do_phys) grid_dependency_wait();
  ENV_STAMP(1);
  uint64_t supp_mask = 0, hit_mask = 0;
  // One jammer (environment.py:248-302, core/jammer.py:73-98): its power-penalty term and what it does to
  // its target radar -- code = (target << 2) | {0 nothing, 1 suppression: val = received power,
  // 2 detected false target: val = 1 - min(pd_f, 0.999999)}.
  auto jammer_eval = [&](int j, double& rp_term, int& code, double& val) {
    const int jr = jbase + 8 * j;
    const int Ti = io.act_d[(int64_t)e * J + j];
    double P = (double)io.act_p[(int64_t)e * J + j];
    P = P < 0.0 ? 0.0 : (P > 1.0 ? 1.0 : P);
    const double pmin = tab(jr + 6), pmax = tab(jr + 7);
    const double range = pmax - pmin;
    const double power = pmin + P * range;
    const double norm = range > 1e-6 ? (power - pmin) / range : 0.0;
    rp_term = T.rp_max + (T.rp_min - T.rp_max) * norm;   // charged even when idle
    code = 0; val = 0.0;
    if (io.jam_power) io.jam_power[(int64_t)j * n + e] = (float)power;
    if (Ti >= 1 && Ti <= 2 * R && power > 0.0) {
      const int tgt = (Ti + 1) / 2 - 1;
      const int rr = rbase + 16 * tgt;
      const double dx = tab(jr + 4) - tab(rr + 10);
      const double dy = tab(jr + 5) - tab(rr + 11);
      const double dist = sqrt(dx * dx + dy * dy);
      if (dist > 1e-6) {
        const double dsq = fmax(1e-9, dist * dist);
        const double den = dsq * tab(jr + 1) * tab(jr + 2) * fmax(1e-9, tab(jr + 3));
        double prj = 0.0;
        if (den > 1e-18) prj = fmax(0.0, (fmax(0.0, power) * tab(jr + 0) * tab(rr + 2)) / den);
        if (Ti & 1) {  // suppression
          code = (tgt << 2) | 1; val = prj;
        } else {       // deception: false target (environment.py:408-437)
          const double pn = tab(rr + 6);
          double snr_f = pn > 1e-18 ? (tab(rr + 8) * prj) / pn : 0.0;
          snr_f = fmax(0.0, snr_f);
          const double pd_f = albersheim(T, snr_f);
          const float u = io.noise ? io.noise[(int64_t)e * (RK + J) + RK + j]
                                   : philox_uniform(io.seed, kStreamEnvNoise, (uint32_t)e, (uint32_t)step, (uint32_t)(RK + j));
          if ((double)u <= pd_f) { code = (tgt << 2) | 2; val = 1.0 - fmin(pd_f, 0.999999); }
        }
      }
    }
  };
  // ... applied in jammer order (the sums and products of the sequential code, term for term)
  auto jammer_apply = [&](double rp_term, int code, double val) {
    r_p += rp_term;
    const int tgt = code >> 2;
    if ((code & 3) == 1) { prjs[tgt * BS + tid] += val; supp_mask |= (1ull << tgt); }
    else if ((code & 3) == 2) { prod[tgt * BS + tid] *= val; hit_mask |= (1ull << tgt); }
  };
  double* jr_rp = tabs + (size_t)(staged ? n_tab_rows : 0) * BS;     // kMode 2: [J][BS] records handed between the workers
  double* jr_val = jr_rp + (size_t)J * BS;
  int* jr_code = reinterpret_cast<int*>(jr_val + (size_t)J * BS);
  if (phys) {
    for (int r = 0; r < R; ++r) { prjs[r * BS + tid] = 0.0; prod[r * BS + tid] = 1.0; }
    if (wk == 0)
      for (int k = 0; k < K; ++k) pnet[k * BS + tid] = 1.0;
    step = io.step_count[e] + 1;  // environment.py:235
#if defined(MACJD_TC_PROFILE) && !defined(MACJD_TEST_HOST_EMULATION)
    if (step == -12345) prjs[tid] = 1.0;   // (keeps the load ahead of the stamp)
#endif
    ENV_STAMP(2);
    if (kMode == 2) {               // each worker evaluates every other jammer
      for (int j = wk; j < J; j += NW) {
        double rp_term, val; int code;
        jammer_eval(j, rp_term, code, val);
        jr_rp[j * BS + tid] = rp_term; jr_val[j * BS + tid] = val; jr_code[j * BS + tid] = code;
      }
    } else {
      for (int j = 0; j < J; ++j) {
        double rp_term, val; int code;
        jammer_eval(j, rp_term, code, val);
        jammer_apply(rp_term, code, val);
      }
    }
  }
#ifndef MACJD_TEST_HOST_EMULATION
  if (kMode == 2 && a.physics && do_phys) asm volatile("bar.sync 2, %0;\n" ::"r"(2 * BS) : "memory");   // jammer records complete
#endif
  if (phys) {
    if (kMode == 2)
      for (int j = 0; j < J; ++j) jammer_apply(jr_rp[j * BS + tid], jr_code[j * BS + tid], jr_val[j * BS + tid]);

    ENV_STAMP(3);
    // ---- radar x target loop (environment.py:316-349, 359-366, 385-398)
    const double four_pi3 = (4.0 * 3.141592653589793) * (4.0 * 3.141592653589793) * (4.0 * 3.141592653589793);
    for (int r = wk; r < R; r += NW) {
      const int rr = rbase + 16 * r;
      const double pt = tab(rr + 0), gt = tab(rr + 1), gr = tab(rr + 2);
      const double lam = tab(rr + 3), loss = tab(rr + 4), latm = tab(rr + 5);
      const double pn = tab(rr + 6), Ga = tab(rr + 7), D = tab(rr + 8);
      const double rx = tab(rr + 10), ry = tab(rr + 11);
      const double jam = D * prjs[r * BS + tid];
      const double den1 = jam + pn;
      const double num0 = pt * gt * gr * (lam * lam);
      bool tracked = false;
      double red = 0.0;
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
        const double snr1 = den1 > 1e-18 ? sig / den1 : 0.0;
        const double pd = albersheim(T, snr1);
        const int slot = r * K + k;
        const float u = io.noise ? io.noise[(int64_t)e * (RK + J) + slot]
                                 : philox_uniform(io.seed, kStreamEnvNoise, (uint32_t)e, (uint32_t)step, (uint32_t)slot);
        const bool det = (double)u <= pd;
        tracked |= det;
        // P_d without jamming only matters for radars hit by suppression this step (r_j)
        if ((supp_mask >> r) & 1ull) red += fmax(0.0, albersheim(T, snr0) - pd);
        if (kMode == 2) pdv[slot * BS + tid] = pd;
        else pnet[k * BS + tid] *= (1.0 - pd);
        const int64_t o = (int64_t)slot * n + e;
        if (io.pd) io.pd[o] = (float)pd;
        if (io.detected) io.detected[o] = det ? 1 : 0;
        if (io.snr0) io.snr0[o] = (float)snr0;
        if (io.snr1) io.snr1[o] = (float)fmax(0.0, snr1);
        if (io.jsr_db) io.jsr_db[o] = 10.0f * log10f((float)(jam / sig));   // float32 output of an extension: float log
      }
      if (io.tracking) io.tracking[(int64_t)r * n + e] = tracked ? 1 : 0;
      // memoryless TRACK state (core/radar.py:90-117) -> r_d; suppression / deception terms of r_j
      const double rd_term = tracked ? fmin(fmax(-tab(rr + 9), T.rd_min), T.rd_max) : 0.0;
      const bool supp = (supp_mask >> r) & 1ull, hit = (hit_mask >> r) & 1ull;
      if (kMode == 2) {
        rad[(0 * R + r) * BS + tid] = rd_term;
        rad[(1 * R + r) * BS + tid] = supp ? red : 0.0;
        rad[(2 * R + r) * BS + tid] = hit ? 1.0 - prod[r * BS + tid] : 0.0;
      } else {
        if (tracked) r_d += rd_term;
        if (supp) r_j_supp += red;
        if (hit) r_j_dec += 1.0 - prod[r * BS + tid];
      }
    }
  }
#ifndef MACJD_TEST_HOST_EMULATION
  ENV_STAMP(4);
  if (kMode == 2 && a.physics && do_phys) asm volatile("bar.sync 2, %0;\n" ::"r"(2 * BS) : "memory");   // both physics sets
#endif
  ENV_STAMP(5);
  if (phys && wk == 0) {
    if (kMode == 2) {
      // the sequential kernel's sums, in its order (a term that kernel skips is an exact + 0.0 here)
      for (int r = 0; r < R; ++r) {
        r_d += rad[(0 * R + r) * BS + tid];
        r_j_supp += rad[(1 * R + r) * BS + tid];
        r_j_dec += rad[(2 * R + r) * BS + tid];
        for (int k = 0; k < K; ++k) pnet[k * BS + tid] *= (1.0 - pdv[(r * K + k) * BS + tid]);
      }
    }
    if (io.pd_net)
      for (int k = 0; k < K; ++k) io.pd_net[(int64_t)k * n + e] = (float)(1.0 - pnet[k * BS + tid]);

    const double r_j = r_j_supp + r_j_dec;
    const double reward = r_d + r_p + r_j;                 // environment.py:457
    const bool term = step >= T.episode_limit;             // environment.py:460
    io.reward[e] = (float)reward;
    if (io.reward64) io.reward64[e] = reward;
    if (io.r_d) io.r_d[e] = (float)r_d;
    if (io.r_p) io.r_p[e] = (float)r_p;
    if (io.r_j) io.r_j[e] = (float)r_j;
    if (io.terminated) io.terminated[e] = term ? 1 : 0;
    io.step_count[e] = (term && io.auto_reset) ? 0 : step;
    ENV_STAMP(6);
  } else if (live && !a.physics && do_phys) {
    io.step_count[e] = 0;                                  // environment.py:203
  }

  // ---- static views (environment.py:479-551): stage one state row per thread
  const bool want_views = (io.state != nullptr) || (io.obs != nullptr);
  if (want_views && do_views) {
    if (live) {
      float* row = stage + (size_t)tid * a.stage_ld;
      const int per = 6 + T.n_types;
      for (int r = 0; r < R; ++r) {
        const int rr = rbase + 16 * r, o = r * per;
        row[o + 0] = (float)env_tab(col, rs, rr + 0);          // pt
        row[o + 1] = (float)env_tab(col, rs, rr + 12);         // theta_m
        row[o + 2] = (float)env_tab(col, rs, rr + 14);         // t_s
        const int ty = (int)env_tab(col, rs, rr + 15);
        for (int c = 0; c < T.n_types; ++c) row[o + 3 + c] = (c == ty) ? 1.0f : 0.0f;
        row[o + 3 + T.n_types] = (float)env_tab(col, rs, rr + 13);  // theta_a
        row[o + 4 + T.n_types] = (float)env_tab(col, rs, rr + 10);
        row[o + 5 + T.n_types] = (float)env_tab(col, rs, rr + 11);
      }
      for (int j = 0; j < J; ++j) {
        row[R * per + 2 * j] = (float)env_tab(col, rs, jbase + 8 * j + 4);
        row[R * per + 2 * j + 1] = (float)env_tab(col, rs, jbase + 8 * j + 5);
      }
    }
#ifndef MACJD_TEST_HOST_EMULATION
    if (kMode != 0) asm volatile("bar.sync 1, %0;\n" ::"r"(BS) : "memory");   // the view warps only
    else __syncthreads();
#else
    __syncthreads();
#endif
    if (kMode != 0) grid_dependency_wait();      // view warps: staged from the tables, nothing written yet
    const int valid = min(BS, a.env_end - e0);
    if ((S & 3) == 0 && (a.stage_ld & 3) == 0) {
      // rows are 16-byte multiples: coalesced float4 stores
      const int S4 = S >> 2, ld4 = a.stage_ld >> 2;
      const float4* st4 = reinterpret_cast<const float4*>(stage);
      if (io.state) {
        float4* dst = reinterpret_cast<float4*>(io.state + (int64_t)e0 * S);
        for (int v = tid; v < valid * S4; v += BS) { const int el = a.magic_s4 ? (int)__umulhi((uint32_t)v, a.magic_s4) : v / S4; dst[v] = st4[el * ld4 + (v - el * S4)]; }
      }
      if (io.obs) {
        const int JS4 = J * S4;
        float4* dst = reinterpret_cast<float4*>(io.obs + (int64_t)e0 * J * S);
        for (int v = tid; v < valid * JS4; v += BS) {
          const int el = a.magic_js4 ? (int)__umulhi((uint32_t)v, a.magic_js4) : v / JS4;
          const int w = v - el * JS4;                                    // position inside the env's J rows
          dst[v] = st4[el * ld4 + (w - (a.magic_s4 ? (int)__umulhi((uint32_t)w, a.magic_s4) : w / S4) * S4)];
        }
      }
    } else {
      if (io.state) {
        float* dst = io.state + (int64_t)e0 * S;
        for (int v = tid; v < valid * S; v += BS) { const int el = v / S; dst[v] = stage[el * a.stage_ld + (v - el * S)]; }
      }
      if (io.obs) {
        const int JS = J * S;
        float* dst = io.obs + (int64_t)e0 * JS;
        for (int v = tid; v < valid * JS; v += BS) { const int el = v / JS; dst[v] = stage[el * a.stage_ld + (v % S)]; }
      }
    }
  }
  if (kMode != 0 && do_views && !want_views) grid_dependency_wait();
  if (io.avail && do_views) {  // all actions always available (environment.py:539-551)
    const int valid = min(BS, a.env_end - e0);
    const int64_t base = (int64_t)e0 * J * A;
    const int total = valid * J * A;
    for (int v = tid; v < total; v += BS) io.avail[base + v] = 1;
  }
}

// tab_rows >= 0: the small-batch kernels (kMode 1 / 2), which also hold the handed-over per-radar terms
// (laid out for both modes) and, when launched behind the agent step, a copy of the scenario columns
inline size_t env_smem_bytes(int R, int K, int stage_ld, int bs, int workers = 1, int tab_rows = -1, int n_jammers = 0) {
  return (size_t)(workers * 2 * R + K) * bs * sizeof(double) + (((size_t)bs * stage_ld + 1) / 2) * sizeof(double) +
         (tab_rows >= 0 ? (size_t)(3 * R + R * K + tab_rows) * bs * sizeof(double) : 0) +
         (workers > 1 ? (size_t)3 * n_jammers * bs * sizeof(double) : 0);     // jammer records handed between the workers
}

// argument checks shared by both step kernels
inline int env_check_args(const macjd_ctx* ctx, const macjd_env_tables* tab, const macjd_env_io* io, int physics) {
  if (!ctx || !tab || !io) return MACJD_ERR_INVALID_ARG;
  if (tab->n_envs < 0 || tab->n_jammers < 1 || tab->n_radars < 1 || tab->n_targets < 1 || tab->n_types < 1)
    return MACJD_ERR_INVALID_ARG;
  if (tab->n_radars > 64 || tab->row_stride > 0x7fffffffll || tab->row_stride < 0) return MACJD_ERR_UNSUPPORTED;
  if (!tab->data || !io->step_count) return MACJD_ERR_INVALID_ARG;
  if (physics && (!io->act_d || !io->act_p || !io->reward)) return MACJD_ERR_INVALID_ARG;
  if (io->env_begin < 0 || io->env_count < 0 || (int64_t)io->env_begin + io->env_count > tab->n_envs) return MACJD_ERR_INVALID_ARG;
  return MACJD_OK;
}

// the kernel on the RAW scenario tables (callers that did not run macjd_env_prepare)
inline int env_launch_raw(const macjd_ctx* ctx, const macjd_env_tables* tab, const macjd_env_io* io, int physics) {
  const int n_step = io->env_count > 0 ? io->env_count : tab->n_envs - io->env_begin;
  if (n_step == 0) return MACJD_OK;
  EnvKernelArgs a;
  a.env_begin = io->env_begin;
  a.env_end = io->env_begin + n_step;
  a.tab = *tab;
  a.io = *io;
  a.state_dim = tab->n_radars * (6 + tab->n_types) + 2 * tab->n_jammers;
  a.n_actions = 2 * tab->n_radars + 1;
  // stage rows: keep 16-byte multiples when S allows float4 stores (pad by 4 -> rows shift
  // by 4 banks, conflict-free for the row-per-thread writes), else pad to odd.
  a.stage_ld = (a.state_dim % 4 == 0) ? a.state_dim + 4 : (a.state_dim | 1);
  a.physics = physics;
  // v / d == umulhi(v, ceil(2^32 / d)) whenever v * d < 2^32 (v < 128 * J * S/4 here); 0 = divide
  {
    const uint64_t s4 = (uint64_t)a.state_dim / 4, js4 = (uint64_t)tab->n_jammers * s4;
    const bool ok = s4 >= 2 && 128ull * js4 * js4 < 0x100000000ull;
    a.magic_s4 = ok ? (uint32_t)((0x100000000ull + s4 - 1) / s4) : 0;
    a.magic_js4 = ok ? (uint32_t)((0x100000000ull + js4 - 1) / js4) : 0;
  }
  // latency-bound regime only: with many resident blocks per SM the loads already overlap
  a.prefetch = tab->env_stride != 0 && tab->n_envs <= 16384;   // measured: +15 % at 4096 envs, -6 % at 65536
  int bs = 128;
  while (bs > 32 && env_smem_bytes(tab->n_radars, tab->n_targets, a.stage_ld, bs) > 96 * 1024) bs >>= 1;
  const size_t smem = env_smem_bytes(tab->n_radars, tab->n_targets, a.stage_ld, bs);
  if (smem > 200 * 1024) return MACJD_ERR_UNSUPPORTED;
  if (smem > 48 * 1024) {
    if (cudaFuncSetAttribute(env_step_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return MACJD_ERR_CUDA;
  }
  const int grid = (n_step + bs - 1) / bs;
#ifndef MACJD_TEST_HOST_EMULATION
  // latency-bound regime: physics and views as concurrent chains (measured at 4 096 envs: 22.5 -> 18.4 us),
  // and the radar loop over two workers
  // MACJD_ENV_FOLLOWS_AGENT: the caller launches this step right behind the kernel that writes its actions;
  // only then is there a wait to fill with the table staging (alone, the staging is 6 us of extra latency)
  a.stage_tab = (io->flags & MACJD_ENV_FOLLOWS_AGENT) ? 1 : 0;
  const int tab_rows = a.stage_tab ? 16 * tab->n_radars + 8 * tab->n_jammers + 3 * tab->n_targets : 0;
  const size_t smem1 = env_smem_bytes(tab->n_radars, tab->n_targets, a.stage_ld, bs, 1, tab_rows);
  const size_t smem2 = env_smem_bytes(tab->n_radars, tab->n_targets, a.stage_ld, bs, 2, tab_rows, tab->n_jammers);
  constexpr size_t kSmallBatchSmem = 110 * 1024;           // two blocks per SM: the grid then fits on the SMs an agent kernel leaves idle
  a.split_views = physics && tab->n_envs <= 16384 && bs == 128 && smem1 <= kSmallBatchSmem && (io->state || io->obs || io->avail);
  if (a.split_views) {
    // allowed to start while the preceding kernel (the agent step) drains: launch latency, table staging and
    // view staging overlap it; the kernel waits (grid_dependency_wait) before the first dependent access
    // two physics workers pay off when the lookups go to L2 (a launch on its own: 16.4 -> 15.0 us); with the
    // columns staged in shared memory one worker is faster (48.0 vs 48.6 us per flushed step: no hand-over barriers)
    const bool two = !a.stage_tab && tab->n_radars >= 2 && smem2 <= kSmallBatchSmem;
    const size_t need = two ? smem2 : smem1;
    static PerDeviceMax opted[4];
    const int dev = ctx->device, which = (two ? 2 : 0) + (a.stage_tab ? 1 : 0);
    if (need > 48 * 1024 && !opted[which].covers(dev, need)) {
      const void* fn = two ? (a.stage_tab ? (const void*)env_step_kernel<2, true> : (const void*)env_step_kernel<2, false>)
                           : (a.stage_tab ? (const void*)env_step_kernel<1, true> : (const void*)env_step_kernel<1, false>);
      const cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need);
      if (e != cudaSuccess) return MACJD_ERR_CUDA;
      opted[which].record(dev, need);
    }
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.gridDim = dim3(grid); cfg.stream = (cudaStream_t)ctx->stream; cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.blockDim = dim3((two ? 3 : 2) * bs); cfg.dynamicSmemBytes = need;
    const cudaError_t err = two ? (a.stage_tab ? cudaLaunchKernelEx(&cfg, env_step_kernel<2, true>, a) : cudaLaunchKernelEx(&cfg, env_step_kernel<2, false>, a))
                                : (a.stage_tab ? cudaLaunchKernelEx(&cfg, env_step_kernel<1, true>, a) : cudaLaunchKernelEx(&cfg, env_step_kernel<1, false>, a));
    return err == cudaSuccess ? MACJD_OK : MACJD_ERR_CUDA;
  }
#else
  a.split_views = 0;
#endif
  MACJD_LAUNCH((env_step_kernel<0, false>), grid, bs, smem, (cudaStream_t)ctx->stream, a);
  return MACJD_OK;
}

}  // namespace macjd
