// C ABI of libmacjd_b200.so (declared in include/macjd.h).  Thin: validate, pick the
// device (and put the caller's current device back on return), enqueue the kernels on the
// caller's stream, translate CUDA errors to status codes.  No allocation, no synchronisation.
// Process-wide state is limited to caches of idempotent driver settings (the per-device
// "shared-memory opt-in already requested" sizes: atomics, a lost race repeats a harmless
// call), the lazily created side stream of the grouped host step (mutex) and the pipeline
// watchdog word (macjd_common.cuh: a pinned host int the tensor-core kernels raise instead of
// trapping); entry points are safe to call from several threads on disjoint buffers.
#include "macjd_common.cuh"
#include "env_step2.cuh"
#include "agent_act.cuh"
#include "agent_act_tc2.cuh"
#include "replay.cuh"
#include "learner.cuh"
#include "agent_unroll.cuh"
#include "gru_rec_rows.cuh"
#include "qhead_repack.cuh"
#include "tc05.cuh"

#include <stdio.h>
#include <mutex>
#include <stdlib.h>
#include <string.h>

namespace {

thread_local char g_last_cuda_error[256] = "";

int finish(const macjd_ctx* ctx, int status) {
  (void)ctx;
  if (status != MACJD_OK) return status;
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s", cudaGetErrorString(err));
    return MACJD_ERR_CUDA;
  }
  return MACJD_OK;
}

// Makes ctx->device current for the duration of one entry point and restores the caller's
// device on every return path (the library must not move a PyTorch thread to another GPU).
struct DeviceScope {
  int prev = -1;
  bool switched = false;
  int status = MACJD_OK;
  explicit DeviceScope(const macjd_ctx* ctx) {
    if (!ctx) { status = MACJD_ERR_INVALID_ARG; return; }
    cudaError_t err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != ctx->device) {
      err = cudaSetDevice(ctx->device);
      switched = err == cudaSuccess;
    }
    if (err != cudaSuccess) {
      snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s", cudaGetErrorString(err));
      status = MACJD_ERR_CUDA;
      return;
    }
    // a tensor-core kernel of an EARLIER call gave up waiting on its pipeline (see macjd_common.cuh)
    if (macjd::watchdog_tripped()) {
      snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "tcgen05 pipeline wait timed out in an earlier launch; its outputs are invalid");
      status = MACJD_ERR_CUDA;
    }
  }
  ~DeviceScope() { if (switched) cudaSetDevice(prev); }
  DeviceScope(const DeviceScope&) = delete;
  DeviceScope& operator=(const DeviceScope&) = delete;
};
// (nested entry points -- the host-buffer calls invoke macjd_agent_forward / macjd_env_step -- find the
// device already current and switch nothing)
#define MACJD_ENTER(ctx)            \
  DeviceScope macjd_scope_(ctx);    \
  if (macjd_scope_.status != MACJD_OK) return macjd_scope_.status

}  // namespace

extern "C" {

const char* macjd_status_string(int status) {
  switch (status) {
    case MACJD_OK: return "ok";
    case MACJD_ERR_INVALID_ARG: return "invalid argument";
    case MACJD_ERR_UNSUPPORTED: return "unsupported dimensions";
    case MACJD_ERR_CUDA: return "CUDA error";
    case MACJD_ERR_WORKSPACE: return "workspace too small";
    default: return "unknown status";
  }
}

const char* macjd_last_cuda_error(void) { return g_last_cuda_error; }

int macjd_abi_version(void) { return MACJD_ABI_VERSION; }

size_t macjd_abi_sizeof(int which) {
  switch (which) {
    case 0: return sizeof(macjd_ctx);
    case 1: return sizeof(macjd_env_tables);
    case 2: return sizeof(macjd_env_io);
    case 3: return sizeof(macjd_agent_weights);
    case 4: return sizeof(macjd_agent_io);
    case 5: return sizeof(macjd_copy_desc);
    case 6: return sizeof(macjd_mixer_dims);
    case 7: return sizeof(macjd_mixer_params);
    case 8: return sizeof(macjd_qhead_dims);
    case 9: return sizeof(macjd_opt_tensors);
    case 10: return sizeof(macjd_act_host);
    case 11: return sizeof(macjd_env_host);
    default: return 0;
  }
}

int macjd_env_step(const macjd_ctx* ctx, const macjd_env_tables* tab, const macjd_env_io* io) {
  MACJD_ENTER(ctx);
  return finish(ctx, macjd::env_launch(ctx, tab, io, /*physics=*/1));
}

int macjd_env_reset(const macjd_ctx* ctx, const macjd_env_tables* tab, const macjd_env_io* io) {
  MACJD_ENTER(ctx);
  return finish(ctx, macjd::env_launch(ctx, tab, io, /*physics=*/0));
}

size_t macjd_env_derived_bytes(const macjd_env_tables* tab) {
  if (!tab || tab->n_envs < 0 || tab->n_jammers < 1 || tab->n_radars < 1 || tab->n_targets < 1 || tab->n_types < 1) return 0;
  return macjd::env_derived_bytes(*tab);
}

int macjd_env_prepare(const macjd_ctx* ctx, const macjd_env_tables* tab, void* derived) {
  MACJD_ENTER(ctx);
  return finish(ctx, macjd::env_prepare_launch(ctx, tab, derived));
}

int macjd_agent_forward(const macjd_ctx* ctx, const macjd_agent_weights* w, const macjd_agent_io* io) {
  MACJD_ENTER(ctx);
  if (!w || !io) return MACJD_ERR_INVALID_ARG;
  if (io->path < 0 || io->path > 3 || io->part < 0 || io->part > 4) return MACJD_ERR_INVALID_ARG;
  if (io->n_rows == 0 && io->n_steps >= 1) return MACJD_OK;      // an empty batch is a no-op (its buffers have no address)
  if (io->part == 2 && (io->n_steps != 1 || !io->hidden)) return MACJD_ERR_INVALID_ARG;
  if (io->part == 3 && (io->n_steps != 1 || !io->gate_x)) return MACJD_ERR_INVALID_ARG;
  if (io->part == 4 && !io->gate_x) return MACJD_ERR_INVALID_ARG;
  // the recurrence of a time-unrolled pass over FEW rows (a learner batch): rows split over CTAs that keep
  // rnn.weight_hh in shared memory (gru_rec_rows.cuh) instead of one CTA pair working through every step
  if (macjd::rec_rows_supported(*w, *io)) return finish(ctx, macjd::rec_rows_launch(ctx, *w, *io));
#ifndef MACJD_TEST_HOST_EMULATION
  if (io->path != 1 && macjd::tc::agent_tc_supported(*w)) {
    if (io->n_rows < 0 || io->n_steps < 1 || (!io->obs && io->part != 4)) return MACJD_ERR_INVALID_ARG;
    if (io->n_rows == 0) return MACJD_OK;
    macjd::AgentArgs a;
    a.w = *w;
    a.io = *io;
    // path 3 / auto: the CTA-pair kernel (cta_group::2, 128 rows per pair) -- also for <= 64 rows (a pair MMA
    // of M = 128, N = 128 takes 37 cycles against 68 for a single-CTA M = 64 one, and each CTA streams half
    // the weights).  path 2 named round 1's single-CTA kernel, which the pair kernel superseded: unsupported.
    if (io->path == 2) return MACJD_ERR_UNSUPPORTED;
    const bool pair_ok = macjd::tc::agent_tc2_supported(*w);
    if (pair_ok) return finish(ctx, macjd::tc::agent_tc2_launch(ctx, a));
    if (io->path == 3 || io->part != 0) return MACJD_ERR_UNSUPPORTED;
    // (auto with dims the pair kernel does not take: the FP32 SIMT kernel below)
  }
#endif
  if (io->path >= 2 || io->part != 0) return MACJD_ERR_UNSUPPORTED;
  return finish(ctx, macjd::agent_launch(ctx, w, io));
}

// One timestep on the device: agent step, then the env step on the actions it chose.  One launch when the CTA-pair
// kernel can run both (macjd_rollout_fused_supported), else the two kernels back to back (the env step as a
// programmatic dependent of the agent step).  Same results either way.
int macjd_rollout_step(const macjd_ctx* ctx, const macjd_agent_weights* w, const macjd_agent_io* aio,
                       const macjd_env_tables* tab, const macjd_env_io* eio) {
  MACJD_ENTER(ctx);
  if (!w || !aio || !tab || !eio) return MACJD_ERR_INVALID_ARG;
  if (aio->n_steps != 1 || aio->part != 0 || !aio->actions || !aio->power) return MACJD_ERR_INVALID_ARG;
  if ((int64_t)aio->n_rows != (int64_t)tab->n_envs * tab->n_jammers) return MACJD_ERR_INVALID_ARG;
  macjd_env_io ke = *eio;
  ke.act_d = aio->actions;
  ke.act_p = aio->power;
  int st = macjd::env_check_args(ctx, tab, &ke, /*physics=*/1);
  if (st != MACJD_OK) return st;
  if (aio->n_rows == 0) return MACJD_OK;
#ifndef MACJD_TEST_HOST_EMULATION
  if (aio->path != 1 && ke.env_begin == 0 && ke.env_count == 0 && macjd::tc::agent_tc_supported(*w) &&
      macjd::tc::agent_tc2_fuse_supported(*w, *tab) && macjd::tc::agent_tc2_fuse_profitable(*tab)) {
    if (aio->n_rows < 0 || !aio->obs) return MACJD_ERR_INVALID_ARG;
    macjd::AgentArgs a;
    a.w = *w;
    a.io = *aio;
    const macjd::Env2Args e = macjd::env2_args(tab, &ke, 1);
    return finish(ctx, macjd::tc::agent_tc2_launch(ctx, a, &e));
  }
#endif
  st = macjd_agent_forward(ctx, w, aio);
  if (st != MACJD_OK) return st;
  ke.flags |= MACJD_ENV_FOLLOWS_AGENT;
  return macjd_env_step(ctx, tab, &ke);
}

// Several timesteps of the rollout loop.  Fused: ONE launch (each CTA pair keeps its rows' recurrent state in shared
// memory, loops over the timesteps and runs its envs' steps itself).  Otherwise: macjd_rollout_step per timestep.
int macjd_rollout_steps(const macjd_ctx* ctx, const macjd_agent_weights* w, const macjd_agent_io* aio,
                        const macjd_env_tables* tab, const macjd_env_io* eio) {
  MACJD_ENTER(ctx);
  if (!w || !aio || !tab || !eio) return MACJD_ERR_INVALID_ARG;
  const int T = aio->n_steps;
  if (T < 1 || aio->part != 0 || !aio->actions || !aio->power || !aio->obs) return MACJD_ERR_INVALID_ARG;
  if ((int64_t)aio->n_rows != (int64_t)tab->n_envs * tab->n_jammers) return MACJD_ERR_INVALID_ARG;
  if (T == 1) return macjd_rollout_step(ctx, w, aio, tab, eio);
  const int64_t M = aio->n_rows, n = tab->n_envs, J = tab->n_jammers, O = w->obs_dim, A = w->n_actions, H = w->hidden;
  const int64_t S = (int64_t)tab->n_radars * (6 + tab->n_types) + 2 * J;
  // the agent of step t + 1 reads what the env of step t wrote: the buffers must be the time-major trajectory
  const int64_t og = aio->obs_group > 1 ? aio->obs_group : 1;
  if (M % og != 0 || O != S) return MACJD_ERR_INVALID_ARG;
  if (og == 1 ? eio->obs != aio->obs + M * O : (og != J || eio->state != aio->obs + n * O)) return MACJD_ERR_INVALID_ARG;
  if (aio->avail && eio->avail != aio->avail + M * A) return MACJD_ERR_INVALID_ARG;
  if (aio->u_eps || aio->rand_actions || eio->noise || aio->rng_step_dev) return MACJD_ERR_UNSUPPORTED;   // (injected draws: step by step)
  macjd_env_io ke = *eio;
  ke.act_d = aio->actions;
  ke.act_p = aio->power;
  int st = macjd::env_check_args(ctx, tab, &ke, /*physics=*/1);
  if (st != MACJD_OK) return st;
  if (M == 0) return MACJD_OK;
#ifndef MACJD_TEST_HOST_EMULATION
  if (aio->path != 1 && ke.env_begin == 0 && ke.env_count == 0 && macjd::tc::agent_tc_supported(*w) &&
      macjd::tc::agent_tc2_fuse_supported(*w, *tab) && macjd::tc::agent_tc2_fuse_profitable(*tab)) {
    macjd::AgentArgs a;
    a.w = *w;
    a.io = *aio;
    const macjd::Env2Args e = macjd::env2_args(tab, &ke, 1);
    return finish(ctx, macjd::tc::agent_tc2_launch(ctx, a, &e));
  }
#endif
  // step by step: the same pointers advanced by one timestep each
  for (int t = 0; t < T && st == MACJD_OK; ++t) {
    macjd_agent_io ka = *aio;
    ka.n_steps = 1;
    ka.obs = aio->obs + t * (M / og) * O;
    if (aio->avail) ka.avail = aio->avail + t * M * A;
    ka.actions = aio->actions + t * M;
    ka.power = aio->power + t * M;
    if (aio->q_chosen) ka.q_chosen = aio->q_chosen + t * M;
    if (aio->hidden_seq) ka.hidden_seq = aio->hidden_seq + t * M * H;
    if (aio->q_all) ka.q_all = aio->q_all + t * M * A;
    if (aio->params_all) ka.params_all = aio->params_all + t * M * A;
    if (aio->greedy) ka.greedy = aio->greedy + t * M;
    if (aio->sel_actions) ka.sel_actions = aio->sel_actions + t * M;
    if (aio->q_sel) ka.q_sel = aio->q_sel + t * M;
    if (aio->actions_mirror) ka.actions_mirror = aio->actions_mirror + t * M;
    if (aio->power_mirror) ka.power_mirror = aio->power_mirror + t * M;
    if (aio->epsilon_dev) ka.epsilon_dev = aio->epsilon_dev + t;
    ka.rng_step = aio->rng_step + (uint32_t)t;
    // the recurrent state: step 0 as given; later steps continue from the previous step's record (or in place)
    if (t > 0) {
      ka.hidden_zero_init = 0;
      ka.hidden_in = aio->hidden_seq ? aio->hidden_seq + (t - 1) * M * H : nullptr;
    }
    if (aio->hidden_seq && aio->hidden && t < T - 1) ka.hidden = nullptr;      // only the last step updates it
    if (!aio->hidden_seq && !aio->hidden) return MACJD_ERR_INVALID_ARG;
    macjd_env_io kt = ke;
    kt.reward = ke.reward + t * n;
    if (ke.r_d) kt.r_d = ke.r_d + t * n;
    if (ke.r_p) kt.r_p = ke.r_p + t * n;
    if (ke.r_j) kt.r_j = ke.r_j + t * n;
    if (ke.terminated) kt.terminated = ke.terminated + t * n;
    if (ke.state) kt.state = ke.state + t * n * S;
    if (ke.obs) kt.obs = ke.obs + t * n * J * S;
    if (ke.avail) kt.avail = ke.avail + t * n * J * A;
    st = macjd_rollout_step(ctx, w, &ka, tab, &kt);
  }
  return st;
}

int macjd_rollout_fused_supported(const macjd_agent_weights* w, const macjd_env_tables* tab) {
#ifndef MACJD_TEST_HOST_EMULATION
  return (w && tab && macjd::tc::agent_tc_supported(*w) && macjd::tc::agent_tc2_fuse_supported(*w, *tab) &&
          macjd::tc::agent_tc2_fuse_profitable(*tab)) ? 1 : 0;
#else
  (void)w; (void)tab;
  return 0;
#endif
}

size_t macjd_agent_unroll_workspace_floats(const macjd_agent_weights* w, int32_t n_rows, int32_t n_steps) {
  if (!w || n_rows < 0 || n_steps < 1) return 0;
  return macjd::unroll_ws_layout(*w, n_rows, n_steps, nullptr).total;
}

int macjd_agent_unroll(const macjd_ctx* ctx, const macjd_agent_weights* w, const macjd_agent_io* io, float* workspace,
                       size_t workspace_floats) {
  MACJD_ENTER(ctx);
  if (!w || !io) return MACJD_ERR_INVALID_ARG;
  return finish(ctx, macjd::agent_unroll_gemm(ctx, *w, *io, workspace, workspace_floats));
}

int macjd_agent_pair_supported(const macjd_agent_weights* w) {
#ifndef MACJD_TEST_HOST_EMULATION
  return (w && macjd::tc::agent_tc2_supported(*w)) ? 1 : 0;
#else
  (void)w;
  return 0;
#endif
}

namespace {
// copies on the launch stream; `what` only labels the error string
int copy_async(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, const macjd_ctx* ctx) {
  if (bytes == 0) return MACJD_OK;
  const cudaError_t err = cudaMemcpyAsync(dst, src, bytes, kind, (cudaStream_t)ctx->stream);
  if (err != cudaSuccess) {
    snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s", cudaGetErrorString(err));
    return MACJD_ERR_CUDA;
  }
  return MACJD_OK;
}
int drain(const macjd_ctx* ctx) {
  const cudaError_t err = cudaStreamSynchronize((cudaStream_t)ctx->stream);
  if (err != cudaSuccess) {
    snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s", cudaGetErrorString(err));
    return MACJD_ERR_CUDA;
  }
  return MACJD_OK;
}

// Host buffers in page-locked memory can be handed to the kernels as they are (the SMs read / write them
// over PCIe: no separate copy operation and none of its launch latency).  Measured on B200 / PCIe 5
// (tools/e2e_host.py, 4096 envs): kernel WRITES to host memory beat a copy-engine transfer even at 786 KB
// (env step + D2H 62 -> 50 us); kernel READS of the 786 KB of observations lost 8 us against the copy
// engine while every thread fetched its own row with 4-byte loads, and win 4 us since the pair kernel
// reads the block in memory order (512 contiguous bytes per warp, requested at kernel entry).  The read
// limit is set just above that size; larger inputs and pageable memory go through the copy engines.
// MACJD_DIRECT_HOST_READ_BYTES / _WRITE_BYTES override the limits (0 = never).
size_t direct_host_limit(bool write) {
  static const size_t lim[2] = {
      [] { const char* e = getenv("MACJD_DIRECT_HOST_READ_BYTES"); return e ? (size_t)strtoull(e, nullptr, 10) : (size_t)1 << 20; }(),
      [] { const char* e = getenv("MACJD_DIRECT_HOST_WRITE_BYTES"); return e ? (size_t)strtoull(e, nullptr, 10) : (size_t)8 << 20; }()};
  return lim[write ? 1 : 0];
}
extern "C++" {
template <typename T>
T* device_alias(T* host_ptr, size_t bytes, bool write, uint32_t flags) {
#ifdef MACJD_TEST_HOST_EMULATION
  (void)bytes; (void)write; (void)flags;
  return host_ptr;                     // emulation: "device" memory is host memory
#else
  if (!host_ptr || bytes > direct_host_limit(write)) return nullptr;
  if (flags & MACJD_HOST_PINNED) return host_ptr;      // caller-vouched: unified addressing, device alias == host pointer
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, host_ptr) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  if (attr.type != cudaMemoryTypeHost || !attr.devicePointer) return nullptr;
  return reinterpret_cast<T*>(attr.devicePointer);
#endif
}
}  // extern "C++"

}  // namespace

int macjd_agent_act_host(const macjd_ctx* ctx, const macjd_agent_weights* w, const macjd_agent_io* io,
                         const macjd_act_host* host) {
  MACJD_ENTER(ctx);
  int st = MACJD_OK;
  if (!w || !io || !host) return MACJD_ERR_INVALID_ARG;
  if (io->n_steps != 1 || io->n_rows < 0) return MACJD_ERR_INVALID_ARG;
  if (!host->obs || !host->actions || !host->power || !io->obs || !io->actions || !io->power) return MACJD_ERR_INVALID_ARG;
  if (host->avail && !io->avail) return MACJD_ERR_INVALID_ARG;
  const size_t M = (size_t)io->n_rows;
  if (M == 0) return MACJD_OK;
  macjd_agent_io k = *io;
  const size_t og = io->obs_group > 1 ? (size_t)io->obs_group : 1;
  if (M % og != 0) return MACJD_ERR_INVALID_ARG;
  const size_t obs_bytes = (M / og) * w->obs_dim * sizeof(float), avail_bytes = M * w->n_actions;
  if (const float* d = device_alias(host->obs, obs_bytes, false, host->flags)) k.obs = d;
  else st = copy_async(const_cast<float*>(io->obs), host->obs, obs_bytes, cudaMemcpyHostToDevice, ctx);
  if (st == MACJD_OK && host->avail) {
    if (const uint8_t* d = device_alias(host->avail, avail_bytes, false, host->flags)) k.avail = d;
    else st = copy_async(const_cast<uint8_t*>(io->avail), host->avail, avail_bytes, cudaMemcpyHostToDevice, ctx);
  }
  if (st != MACJD_OK) return st;
  int32_t* d_act = device_alias(host->actions, M * sizeof(int32_t), true, host->flags);
  float* d_pow = device_alias(host->power, M * sizeof(float), true, host->flags);
  float* d_q = (host->q_chosen && io->q_chosen) ? device_alias(host->q_chosen, M * sizeof(float), true, host->flags) : nullptr;
  if (d_act) k.actions = d_act;
  if (d_pow) k.power = d_pow;
  if (d_q) k.q_chosen = d_q;
  st = macjd_agent_forward(ctx, w, &k);
  if (st != MACJD_OK) return st;
  if (!d_act) st = copy_async(host->actions, io->actions, M * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx);
  if (st == MACJD_OK && !d_pow) st = copy_async(host->power, io->power, M * sizeof(float), cudaMemcpyDeviceToHost, ctx);
  if (st == MACJD_OK && !d_q && host->q_chosen && io->q_chosen)
    st = copy_async(host->q_chosen, io->q_chosen, M * sizeof(float), cudaMemcpyDeviceToHost, ctx);
  if (st != MACJD_OK) return st;
  return drain(ctx);
}

int macjd_env_step_host(const macjd_ctx* ctx, const macjd_env_tables* tab, const macjd_env_io* io,
                        const macjd_env_host* host) {
  MACJD_ENTER(ctx);
  int st = MACJD_OK;
  if (!tab || !io || !host) return MACJD_ERR_INVALID_ARG;
  if (!host->act_d || !host->act_p || !io->act_d || !io->act_p) return MACJD_ERR_INVALID_ARG;
  if ((host->reward && !io->reward) || (host->terminated && !io->terminated) || (host->obs && !io->obs) ||
      (host->state && !io->state))
    return MACJD_ERR_INVALID_ARG;
  if (tab->n_envs <= 0) return tab->n_envs == 0 ? MACJD_OK : MACJD_ERR_INVALID_ARG;
  const size_t n = (size_t)tab->n_envs, J = (size_t)tab->n_jammers;
  const size_t S = (size_t)tab->n_radars * (6 + tab->n_types) + 2 * J;
  macjd_env_io k = *io;
  if (const int32_t* d = device_alias(host->act_d, n * J * sizeof(int32_t), false, host->flags)) k.act_d = d;
  else st = copy_async(const_cast<int32_t*>(io->act_d), host->act_d, n * J * sizeof(int32_t), cudaMemcpyHostToDevice, ctx);
  if (st == MACJD_OK) {
    if (const float* d = device_alias(host->act_p, n * J * sizeof(float), false, host->flags)) k.act_p = d;
    else st = copy_async(const_cast<float*>(io->act_p), host->act_p, n * J * sizeof(float), cudaMemcpyHostToDevice, ctx);
  }
  if (st != MACJD_OK) return st;
  float* d_rew = device_alias(host->reward, n * sizeof(float), true, host->flags);
  uint8_t* d_term = device_alias(host->terminated, n, true, host->flags);
  float* d_obs = device_alias(host->obs, n * J * S * sizeof(float), true, host->flags);
  float* d_state = device_alias(host->state, n * S * sizeof(float), true, host->flags);
  if (d_rew) k.reward = d_rew;
  if (d_term) k.terminated = d_term;
  if (d_obs) k.obs = d_obs;
  if (d_state) k.state = d_state;
  st = macjd_env_step(ctx, tab, &k);
  if (st != MACJD_OK) return st;
  if (host->reward && !d_rew) st = copy_async(host->reward, io->reward, n * sizeof(float), cudaMemcpyDeviceToHost, ctx);
  if (st == MACJD_OK && host->terminated && !d_term) st = copy_async(host->terminated, io->terminated, n, cudaMemcpyDeviceToHost, ctx);
  if (st == MACJD_OK && host->obs && !d_obs) st = copy_async(host->obs, io->obs, n * J * S * sizeof(float), cudaMemcpyDeviceToHost, ctx);
  if (st == MACJD_OK && host->state && !d_state) st = copy_async(host->state, io->state, n * S * sizeof(float), cudaMemcpyDeviceToHost, ctx);
  if (st != MACJD_OK) return st;
  return drain(ctx);
}

namespace {
// Second stream + fork / join events of the grouped host step, one set per device, created on first use.
struct HostGroupStreams { cudaStream_t side = nullptr; cudaEvent_t fork = nullptr, join = nullptr; bool failed = false; };
HostGroupStreams* host_group_streams(int dev) {
  static HostGroupStreams all[64];
  static std::mutex mu;
  if (dev < 0 || dev >= 64) return nullptr;         // (no grouping on devices beyond the table)
  std::lock_guard<std::mutex> lock(mu);
  HostGroupStreams& h = all[dev];
  if (!h.side && !h.failed) {
    if (cudaStreamCreateWithFlags(&h.side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h.join, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      h.failed = true;
    }
  }
  return h.failed ? nullptr : &h;
}
// A batch of at least 2 x this many envs is stepped as two groups on two streams (MACJD_HOST_GROUP_ENVS; 0 = never, the default)
size_t host_group_envs() {
  const char* e = getenv("MACJD_HOST_GROUP_ENVS");      // read per call (tens of ns): tests switch it
  return e ? (size_t)strtoull(e, nullptr, 10) : (size_t)0;
}
extern "C++" {
template <typename T>
T* offset_ptr(T* p, size_t elems) { return p ? p + elems : nullptr; }
}
}  // namespace

// Optional grouping (MACJD_HOST_GROUP_ENVS = g > 0: batches of >= 2 g envs are stepped as two groups of envs on
// two streams, group B's chain issued behind group A's, so that B's observations come in over PCIe while A's
// agent kernel runs and A's results go out while B computes).  Rows, envs and Philox counters keep their
// whole-batch indices (io.rng_row_offset, io.env_begin / env_count): results do not depend on the grouping
// (tests/runner_checks.py: check_fused_host_step).  OFF by default -- measured on B200, 4 096 envs x 2 jammers
// (tools/e2e_host.py): one group 87.0 us per step, two groups 93.6 us: the second pair of launches and the
// fork / join cost more than the overlap returns at this size.
int macjd_rollout_step_host(const macjd_ctx* ctx, const macjd_agent_weights* w, const macjd_agent_io* aio,
                            const macjd_act_host* ahost, const macjd_env_tables* tab, const macjd_env_io* eio,
                            const macjd_env_host* ehost) {
  MACJD_ENTER(ctx);
  int st = MACJD_OK;
  if (!w || !aio || !ahost || !tab || !eio || !ehost) return MACJD_ERR_INVALID_ARG;
  if (aio->n_steps != 1 || aio->n_rows < 0 || tab->n_envs < 0 || tab->n_jammers < 1) return MACJD_ERR_INVALID_ARG;
  if ((int64_t)aio->n_rows != (int64_t)tab->n_envs * tab->n_jammers) return MACJD_ERR_INVALID_ARG;
  if (eio->env_begin != 0 || eio->env_count != 0) return MACJD_ERR_INVALID_ARG;
  if (!ahost->obs || !aio->obs || !aio->actions || !aio->power) return MACJD_ERR_INVALID_ARG;
  if (ahost->avail && !aio->avail) return MACJD_ERR_INVALID_ARG;
  if ((ehost->reward && !eio->reward) || (ehost->terminated && !eio->terminated) || (ehost->obs && !eio->obs) ||
      (ehost->state && !eio->state))
    return MACJD_ERR_INVALID_ARG;
  const size_t M = (size_t)aio->n_rows;
  if (M == 0) return MACJD_OK;
  const size_t n = (size_t)tab->n_envs, J = (size_t)tab->n_jammers;
  const size_t S = (size_t)tab->n_radars * (6 + tab->n_types) + 2 * J;
  const size_t O = (size_t)w->obs_dim, A = (size_t)w->n_actions, H = (size_t)w->hidden;
  // in-place or copy-engine: decided on the whole buffers, so that grouping does not change the transfer mode
  const size_t og = aio->obs_group > 1 ? (size_t)aio->obs_group : 1;
  if (og > 1 && og != J) return MACJD_ERR_INVALID_ARG;         // (one observation row per env, or one per agent)
  const float* a_obs = device_alias(ahost->obs, (M / og) * O * sizeof(float), false, ahost->flags);
  const uint8_t* a_avail = ahost->avail ? device_alias(ahost->avail, M * A, false, ahost->flags) : nullptr;
  float* a_q = (ahost->q_chosen && aio->q_chosen) ? device_alias(ahost->q_chosen, M * sizeof(float), true, ahost->flags) : nullptr;
  int32_t* a_act = device_alias(ahost->actions, M * sizeof(int32_t), true, ahost->flags);
  float* a_pow = device_alias(ahost->power, M * sizeof(float), true, ahost->flags);
  float* e_rew = device_alias(ehost->reward, n * sizeof(float), true, ehost->flags);
  uint8_t* e_term = device_alias(ehost->terminated, n, true, ehost->flags);
  float* e_obs = device_alias(ehost->obs, n * J * S * sizeof(float), true, ehost->flags);
  float* e_state = device_alias(ehost->state, n * S * sizeof(float), true, ehost->flags);

  // groups of envs: [0, n0) on the caller's stream, [n0, n) on the side stream
  size_t n0 = n;
  HostGroupStreams* gs = nullptr;
  if (host_group_envs() > 0 && n >= 2 * host_group_envs() && (gs = host_group_streams(ctx->device)) != nullptr)
    n0 = ((n / 2 + 63) / 64) * 64;
  const int n_groups = n0 < n ? 2 : 1;
  macjd_ctx gctx[2] = {*ctx, *ctx};
  if (n_groups == 2) {
    gctx[1].stream = gs->side;
    if (cudaEventRecord(gs->fork, (cudaStream_t)ctx->stream) != cudaSuccess ||
        cudaStreamWaitEvent(gs->side, gs->fork, 0) != cudaSuccess) {
      snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s", cudaGetErrorString(cudaGetLastError()));
      return MACJD_ERR_CUDA;
    }
  }
  // a failure after the fork still joins and drains both streams before returning
  for (int g = 0; g < n_groups && st == MACJD_OK; ++g) {
    const macjd_ctx* c = &gctx[g];
    const size_t eb = g == 0 ? 0 : n0, ec = g == 0 ? n0 : n - n0;   // envs of this group
    const size_t rb = eb * J, rc = ec * J;                           // agent rows of this group
    // ---- agent: host observations in; the actions stay in the device staging buffers for the env kernel
    macjd_agent_io ka = *aio;
    ka.n_rows = (int32_t)rc;
    ka.rng_row_offset = aio->rng_row_offset + (int32_t)rb;
    ka.obs = a_obs ? a_obs + (rb / og) * O : aio->obs + (rb / og) * O;
    ka.avail = (ahost->avail && a_avail) ? a_avail + rb * A : offset_ptr(aio->avail, rb * A);
    ka.hidden = offset_ptr(aio->hidden, rb * H);
    ka.hidden_in = offset_ptr(aio->hidden_in, rb * H);
    ka.hidden_seq = offset_ptr(aio->hidden_seq, rb * H);
    ka.q_all = offset_ptr(aio->q_all, rb * A);
    ka.params_all = offset_ptr(aio->params_all, rb * A);
    ka.greedy = offset_ptr(aio->greedy, rb);
    ka.sel_actions = offset_ptr(aio->sel_actions, rb);
    ka.q_sel = offset_ptr(aio->q_sel, rb);
    ka.u_eps = offset_ptr(aio->u_eps, rb);
    ka.rand_actions = offset_ptr(aio->rand_actions, rb);
    ka.actions = aio->actions + rb;
    ka.power = aio->power + rb;
    ka.q_chosen = a_q ? a_q + rb : offset_ptr(aio->q_chosen, rb);
    // page-locked action buffers: the agent kernel writes them as well (no copy-engine operation behind the env kernel)
    ka.actions_mirror = a_act ? a_act + rb : nullptr;
    ka.power_mirror = a_pow ? a_pow + rb : nullptr;
    if (!a_obs) st = copy_async(const_cast<float*>(ka.obs), ahost->obs + (rb / og) * O, (rc / og) * O * sizeof(float), cudaMemcpyHostToDevice, c);
    if (st == MACJD_OK && ahost->avail && !a_avail)
      st = copy_async(const_cast<uint8_t*>(ka.avail), ahost->avail + rb * A, rc * A, cudaMemcpyHostToDevice, c);
    if (st != MACJD_OK) break;
    // ---- env: reads the chosen actions where the agent kernel left them
    macjd_env_io ke = *eio;
    ke.act_d = aio->actions;
    ke.act_p = aio->power;
    ke.flags |= MACJD_ENV_FOLLOWS_AGENT;
    ke.env_begin = (int32_t)eb;
    ke.env_count = (int32_t)ec;
    if (e_rew) ke.reward = e_rew;
    if (e_term) ke.terminated = e_term;
    if (e_obs) ke.obs = e_obs;
    if (e_state) ke.state = e_state;
    if (n_groups == 1) {
      ke.env_begin = ke.env_count = 0;
      st = macjd_rollout_step(c, w, &ka, tab, &ke);      // one launch when the pair kernel can run both
    } else {
      st = macjd_agent_forward(c, w, &ka);
      if (st == MACJD_OK) st = macjd_env_step(c, tab, &ke);
    }
  }
  // ---- what the kernels did not write in place (after both chains are issued: copies queue behind their group)
  for (int g = 0; g < n_groups && st == MACJD_OK; ++g) {
    const macjd_ctx* c = &gctx[g];
    const size_t eb = g == 0 ? 0 : n0, ec = g == 0 ? n0 : n - n0, rb = eb * J, rc = ec * J;
    if (ahost->actions && !a_act) st = copy_async(ahost->actions + rb, aio->actions + rb, rc * sizeof(int32_t), cudaMemcpyDeviceToHost, c);
    if (st == MACJD_OK && ahost->power && !a_pow) st = copy_async(ahost->power + rb, aio->power + rb, rc * sizeof(float), cudaMemcpyDeviceToHost, c);
    if (st == MACJD_OK && !a_q && ahost->q_chosen && aio->q_chosen)
      st = copy_async(ahost->q_chosen + rb, aio->q_chosen + rb, rc * sizeof(float), cudaMemcpyDeviceToHost, c);
    if (st == MACJD_OK && ehost->reward && !e_rew) st = copy_async(ehost->reward + eb, eio->reward + eb, ec * sizeof(float), cudaMemcpyDeviceToHost, c);
    if (st == MACJD_OK && ehost->terminated && !e_term) st = copy_async(ehost->terminated + eb, eio->terminated + eb, ec, cudaMemcpyDeviceToHost, c);
    if (st == MACJD_OK && ehost->obs && !e_obs)
      st = copy_async(ehost->obs + eb * J * S, eio->obs + eb * J * S, ec * J * S * sizeof(float), cudaMemcpyDeviceToHost, c);
    if (st == MACJD_OK && ehost->state && !e_state)
      st = copy_async(ehost->state + eb * S, eio->state + eb * S, ec * S * sizeof(float), cudaMemcpyDeviceToHost, c);
  }
  if (n_groups == 2) {
    if (cudaEventRecord(gs->join, gs->side) != cudaSuccess || cudaStreamWaitEvent((cudaStream_t)ctx->stream, gs->join, 0) != cudaSuccess) {
      cudaStreamSynchronize(gs->side);
      if (st == MACJD_OK) {
        snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s", cudaGetErrorString(cudaGetLastError()));
        st = MACJD_ERR_CUDA;
      }
    }
  }
  const int dr = drain(ctx);
  return st != MACJD_OK ? st : dr;
}

int macjd_replay_copy(const macjd_ctx* ctx, const macjd_copy_desc* descs_host, int32_t n_keys, const int32_t* idx,
                      int32_t n_eps, int32_t index_on_src) {
  MACJD_ENTER(ctx);
  return finish(ctx, macjd::replay_copy(ctx, descs_host, n_keys, idx, n_eps, index_on_src));
}

static bool mixer_dims_ok(const macjd_mixer_dims* d) {
  return d && d->n_rows >= 0 && d->state_dim > 0 && d->n_agents > 0 && d->embed_dim > 0 && d->hyper_hidden > 0;
}

size_t macjd_mixer_workspace_floats(const macjd_mixer_dims* dims) {
  if (!mixer_dims_ok(dims)) return 0;
  return macjd::mixer_ws_layout(*dims, nullptr).total;
}

int macjd_mixer_forward(const macjd_ctx* ctx, const macjd_mixer_dims* dims, const macjd_mixer_params* w, const float* q,
                        const float* states, float* q_tot, float* workspace, size_t workspace_floats) {
  MACJD_ENTER(ctx);
  if (!mixer_dims_ok(dims) || !w || !q || !states || !q_tot) return MACJD_ERR_INVALID_ARG;
  return finish(ctx, macjd::mixer_forward((cudaStream_t)ctx->stream, *dims, *w, q, states, q_tot, workspace, workspace_floats));
}

int macjd_mixer_backward(const macjd_ctx* ctx, const macjd_mixer_dims* dims, const macjd_mixer_params* w, const float* q,
                         const float* dq_tot, float* workspace, size_t workspace_floats, const macjd_mixer_params* grads,
                         float* dq) {
  MACJD_ENTER(ctx);
  if (!mixer_dims_ok(dims) || !w || !q || !dq_tot || !grads) return MACJD_ERR_INVALID_ARG;
  return finish(ctx, macjd::mixer_backward((cudaStream_t)ctx->stream, *dims, *w, q, dq_tot, workspace, workspace_floats, *grads, dq));
}

size_t macjd_qhead_scratch_floats(const macjd_qhead_dims* dims) {
  if (!dims || dims->n_rows < 0 || dims->hidden < 1 || dims->n_actions < 1) return 0;
  return macjd::qhead_scratch_floats(*dims);
}

int macjd_qhead_forward(const macjd_ctx* ctx, const macjd_qhead_dims* dims, const macjd_agent_weights* w,
                        const float* hidden, const int32_t* a_d, const float* a_c, float* q, float* hid) {
  MACJD_ENTER(ctx);
  if (!dims || !w || !hidden || !a_d || !a_c || !q || !hid || dims->n_rows < 0) return MACJD_ERR_INVALID_ARG;
  if (dims->hidden != w->hidden || dims->n_actions != w->n_actions) return MACJD_ERR_INVALID_ARG;
  return finish(ctx, macjd::qhead_forward((cudaStream_t)ctx->stream, *dims, *w, hidden, a_d, a_c, q, hid, nullptr, 0));
}

int macjd_qhead_forward_ws(const macjd_ctx* ctx, const macjd_qhead_dims* dims, const macjd_agent_weights* w,
                           const float* hidden, const int32_t* a_d, const float* a_c, float* q, float* hid, float* scratch,
                           size_t scratch_floats) {
  MACJD_ENTER(ctx);
  if (!dims || !w || !hidden || !a_d || !a_c || !q || !hid || dims->n_rows < 0) return MACJD_ERR_INVALID_ARG;
  if (dims->hidden != w->hidden || dims->n_actions != w->n_actions) return MACJD_ERR_INVALID_ARG;
  return finish(ctx, macjd::qhead_forward((cudaStream_t)ctx->stream, *dims, *w, hidden, a_d, a_c, q, hid, scratch, scratch_floats));
}

int macjd_qhead_backward(const macjd_ctx* ctx, const macjd_qhead_dims* dims, const macjd_agent_weights* w,
                         const float* hidden, const int32_t* a_d, const float* a_c, float* hid, const float* dq,
                         float* g_w1, float* g_b1, float* g_w2, float* g_b2, float* scratch, size_t scratch_floats) {
  MACJD_ENTER(ctx);
  if (!dims || !w || !hidden || !a_d || !a_c || !hid || !dq || !g_w1 || !g_b1 || !g_w2 || !g_b2) return MACJD_ERR_INVALID_ARG;
  if (dims->hidden != w->hidden || dims->n_actions != w->n_actions) return MACJD_ERR_INVALID_ARG;
  return finish(ctx, macjd::qhead_backward((cudaStream_t)ctx->stream, *dims, *w, hidden, a_d, a_c, hid, dq, g_w1, g_b1,
                                           g_w2, g_b2, scratch, scratch_floats));
}

int macjd_gather_q(const macjd_ctx* ctx, int32_t n, int32_t n_actions, const float* q_all, const int32_t* idx, float* out) {
  MACJD_ENTER(ctx);
  if (n < 0 || n_actions < 1 || !q_all || !idx || !out) return MACJD_ERR_INVALID_ARG;
  if (n == 0) return MACJD_OK;
  MACJD_LAUNCH(macjd::gather_q_kernel, dim3((n + 255) / 256), dim3(256), 0, (cudaStream_t)ctx->stream, q_all,
               (const int*)idx, (int)n, (int)n_actions, out);
  return finish(ctx, MACJD_OK);
}

size_t macjd_td_scratch_floats(int32_t n_rows) { return n_rows > 0 ? macjd::td_scratch_floats(n_rows) : 0; }

int macjd_td_loss(const macjd_ctx* ctx, int32_t n_rows, const float* q_tot, const float* tq_tot, const float* reward,
                  const uint8_t* terminated, const uint8_t* filled, float gamma, float* dq_tot, float* targets,
                  float* sums, float* scratch, size_t scratch_floats) {
  MACJD_ENTER(ctx);
  if (!q_tot || !tq_tot || !reward || !terminated || !filled || !dq_tot || !sums) return MACJD_ERR_INVALID_ARG;
  return finish(ctx, macjd::td_loss((cudaStream_t)ctx->stream, n_rows, q_tot, tq_tot, reward, terminated, filled, gamma,
                                    dq_tot, targets, sums, scratch, scratch_floats));
}

size_t macjd_opt_scratch_floats(void) { return macjd::opt_scratch_floats(); }

int macjd_clip_adam(const macjd_ctx* ctx, const macjd_opt_tensors* tensors, const float* grad, float* m, float* v,
                    const float* sums, float max_norm, float lr, float beta1, float beta2, float eps, int64_t step,
                    float* scal, float* scratch, size_t scratch_floats) {
  MACJD_ENTER(ctx);
  if (!tensors || !grad || !m || !v || !sums || !scal) return MACJD_ERR_INVALID_ARG;
  return finish(ctx, macjd::clip_adam((cudaStream_t)ctx->stream, *tensors, grad, m, v, sums, max_norm, lr, beta1, beta2,
                                      eps, step, scal, scratch, scratch_floats));
}

int macjd_clip_adam_dev(const macjd_ctx* ctx, const macjd_opt_tensors* tensors, const float* grad, float* m, float* v,
                        const float* sums, float max_norm, float lr, float beta1, float beta2, float eps,
                        const float* bias_corr, float* scal, float* scratch, size_t scratch_floats) {
  MACJD_ENTER(ctx);
  if (!tensors || !grad || !m || !v || !sums || !scal || !bias_corr) return MACJD_ERR_INVALID_ARG;
  return finish(ctx, macjd::clip_adam((cudaStream_t)ctx->stream, *tensors, grad, m, v, sums, max_norm, lr, beta1, beta2,
                                      eps, 0, scal, scratch, scratch_floats, bias_corr));
}

int macjd_qhead_repack(const macjd_ctx* ctx, const macjd_agent_weights* w, const float* w1, const float* b1, const float* w2,
                       const float* b2, float* tc_chunks, int32_t tc_kc, float* tc_q_c, float* tc_w1a, int32_t tc_w1a_stride) {
  MACJD_ENTER(ctx);
  if (!w) return MACJD_ERR_INVALID_ARG;
  macjd::QheadRepackArgs a;
  a.H = w->hidden; a.A = w->n_actions;
  a.w1 = w1; a.b1 = b1; a.w2 = w2; a.b2 = b2;
  // (the packed fields are const for the kernels that READ them; this is the one call that writes them)
  a.wqt = const_cast<float*>(w->wqt); a.bq1 = const_cast<float*>(w->bq1); a.w1a = const_cast<float*>(w->w1a);
  a.w1p = const_cast<float*>(w->w1p); a.w2p = const_cast<float*>(w->w2); a.bq2 = const_cast<float*>(w->bq2);
  a.tc_chunks = tc_chunks; a.kc = tc_kc; a.tc_q_c = tc_q_c; a.tc_w1a = tc_w1a; a.tc_w1a_stride = tc_w1a_stride;
  return finish(ctx, macjd::qhead_repack((cudaStream_t)ctx->stream, a));
}

int macjd_gemm(const macjd_ctx* ctx, int32_t M, int32_t N, int32_t K, const float* A, int32_t lda, int32_t ta, const float* B,
               int32_t ldb, int32_t tb, float* C, int32_t ldc, const float* bias, int32_t act, int32_t accumulate, float* splitk_ws,
               size_t splitk_ws_floats) {
  MACJD_ENTER(ctx);
  if (M < 0 || N < 0 || K < 1 || !A || !B || !C || act < 0 || act > 3 || act == 2) return MACJD_ERR_INVALID_ARG;
  macjd::GemmOpts o;
  o.bias = bias; o.act = act; o.accumulate = accumulate; o.splitk_ws = splitk_ws; o.splitk_ws_floats = splitk_ws_floats;
  // the same workspace lets a tall product pre-split B once per call (K < 2048: split-K never applies there)
  if (K < 2048) { o.bpack_ws = splitk_ws; o.bpack_ws_floats = splitk_ws ? splitk_ws_floats : 0; }
  macjd::gemm((cudaStream_t)ctx->stream, A, lda, ta != 0, B, ldb, tb != 0, C, ldc, M, N, K, o);
  return finish(ctx, MACJD_OK);
}

int macjd_tc_gemm_selftest(const macjd_ctx* ctx, int32_t M, int32_t N, int32_t K, const float* A, const float* B,
                           float* D) {
  MACJD_ENTER(ctx);
#ifdef MACJD_TEST_HOST_EMULATION
  (void)M; (void)N; (void)K; (void)A; (void)B; (void)D;
  return MACJD_ERR_UNSUPPORTED;   // tensor cores cannot be emulated on the host
#else
  // N's bit 30 selects the fragment-layout TMEM read (tcgen05.ld.16x256b) in the epilogue
  return finish(ctx, macjd::tc::tc_gemm_selftest(ctx, M, N & 0xFFFF, K, A, B, D, (N >> 30) & 1));
#endif
}

int macjd_clear_pipeline_fault(void) {
#ifndef MACJD_TEST_HOST_EMULATION
  if (macjd::g_watchdog_seen) *reinterpret_cast<volatile int*>(macjd::g_watchdog_seen) = 0;
#endif
  return MACJD_OK;
}

// k-extent of one packed weight chunk the tcgen05 agent kernel was built for (0: no tensor-core path).
int macjd_agent_tc_chunk_k(void) {
#ifdef MACJD_TEST_HOST_EMULATION
  return 0;
#else
  return macjd::tc::agent_tc_chunk_k();
#endif
}

// ---- development aids (tools/*.py): phase timestamps and tensor-pipe micro-benchmarks.  Not part of the ABI and
// not in the product library: compiled only into the tooling build (-DMACJD_TC_PROFILE or -DMACJD_DEBUG_TOOLS,
// tools/_prof/build_prof.py).
#if defined(MACJD_TC_PROFILE) || defined(MACJD_DEBUG_TOOLS)
// Debug aid (not part of the documented ABI): phase timestamps of the tcgen05 agent kernel when
// the library was compiled with -DMACJD_TC_PROFILE; zeros otherwise.
__attribute__((visibility("default"))) int macjd_debug_tc_profile(unsigned long long* out_host, int n) {
#ifdef MACJD_TEST_HOST_EMULATION
  for (int i = 0; i < n; ++i) out_host[i] = 0;
  return MACJD_OK;
#else
  return macjd::tc::tc_profile_read(out_host, n);
#endif
}

#if defined(MACJD_TC_PROFILE) && !defined(MACJD_TEST_HOST_EMULATION)
// Debug aid: the env kernel's phase stamps (tools/env_phase_profile.py)
__attribute__((visibility("default"))) int macjd_debug_env_profile(unsigned long long* out_host) {
  return cudaMemcpyFromSymbol(out_host, macjd::g_env_prof, sizeof(unsigned long long) * 16) == cudaSuccess ? 0 : -3;
}
#endif

// Debug aid: tcgen05.mma issue / completion rate (cycles) for n back-to-back M x N x 8 TF32 MMAs.
__attribute__((visibility("default"))) int macjd_debug_tc_mma_rate(const macjd_ctx* ctx, int M, int N, int n,
                                                                  unsigned long long* out_dev) {
  MACJD_ENTER(ctx);
#ifdef MACJD_TEST_HOST_EMULATION
  (void)M; (void)N; (void)n; (void)out_dev;
  return MACJD_ERR_UNSUPPORTED;
#else
  return finish(ctx, macjd::tc::tc_mma_rate(ctx, M, N, n, out_dev));
#endif
}

__attribute__((visibility("default"))) int macjd_debug_tc_mma_rate_multi(const macjd_ctx* ctx, int M, int N, int n, int issuers,
                                                                        unsigned long long* out_dev) {
  MACJD_ENTER(ctx);
#ifdef MACJD_TEST_HOST_EMULATION
  (void)M; (void)N; (void)n; (void)issuers; (void)out_dev;
  return MACJD_ERR_UNSUPPORTED;
#else
  return finish(ctx, macjd::tc::tc_mma_rate_multi(ctx, M, N, n, issuers, out_dev));
#endif
}

// Same, for tcgen05.mma.cta_group::2 issued by the leader of a 2-CTA cluster (M = pair rows: 128 or 256).
__attribute__((visibility("default"))) int macjd_debug_tc2_mma_rate(const macjd_ctx* ctx, int M, int N, int n,
                                                                   unsigned long long* out_dev) {
  MACJD_ENTER(ctx);
#ifdef MACJD_TEST_HOST_EMULATION
  (void)M; (void)N; (void)n; (void)out_dev;
  return MACJD_ERR_UNSUPPORTED;
#else
  return finish(ctx, macjd::tc::tc2_mma_rate(ctx, M, N, n, out_dev));
#endif
}

#endif  // tooling build

}  // extern "C"
