// C ABI of libmacjd_b200.so (declared in include/macjd.h).  Thin: validate, pick the
// device, enqueue the kernels on the caller's stream, translate CUDA errors to status
// codes.  No allocation, no synchronisation, no global mutable state.
#include "macjd_common.cuh"
#include "env_step.cuh"
#include "agent_act.cuh"

#include <stdio.h>
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

int enter(const macjd_ctx* ctx) {
  if (!ctx) return MACJD_ERR_INVALID_ARG;
  const cudaError_t err = cudaSetDevice(ctx->device);
  if (err != cudaSuccess) {
    snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s", cudaGetErrorString(err));
    return MACJD_ERR_CUDA;
  }
  return MACJD_OK;
}

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
    default: return 0;
  }
}

int macjd_env_step(const macjd_ctx* ctx, const macjd_env_tables* tab, const macjd_env_io* io) {
  int st = enter(ctx);
  if (st != MACJD_OK) return st;
  return finish(ctx, macjd::env_launch(ctx, tab, io, /*physics=*/1));
}

int macjd_env_reset(const macjd_ctx* ctx, const macjd_env_tables* tab, const macjd_env_io* io) {
  int st = enter(ctx);
  if (st != MACJD_OK) return st;
  return finish(ctx, macjd::env_launch(ctx, tab, io, /*physics=*/0));
}

int macjd_agent_forward(const macjd_ctx* ctx, const macjd_agent_weights* w, const macjd_agent_io* io) {
  int st = enter(ctx);
  if (st != MACJD_OK) return st;
  return finish(ctx, macjd::agent_launch(ctx, w, io));
}

}  // extern "C"
