"""ctypes binding of libmacjd_b200.so (C ABI in include/macjd.h).

There is no CPU fallback: if the CUDA library has not been built, importing a product
class that needs it raises.  Build it with ``python __graft_entry__.py`` (or
``python <package>/csrc/build.py``).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MACJD_LIB_PATH", os.path.join(_HERE, "csrc", "libmacjd_b200.so"))   # override: profiling builds

c_void_p, c_int32, c_int64, c_uint64, c_double, c_float = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double, C.c_float


class Ctx(C.Structure):
    _fields_ = [("device", c_int32), ("reserved", c_int32), ("stream", c_void_p)]


class EnvTables(C.Structure):
    _fields_ = [("n_envs", c_int32), ("n_jammers", c_int32), ("n_radars", c_int32), ("n_targets", c_int32),
                ("n_types", c_int32), ("episode_limit", c_int32),
                ("data", c_void_p), ("row_stride", c_int64), ("env_stride", c_int64),
                ("rd_min", c_double), ("rd_max", c_double), ("rp_min", c_double), ("rp_max", c_double),
                ("alb_a", c_double), ("alb_zoff", c_double), ("alb_den", c_double), ("derived", c_void_p)]


class EnvIO(C.Structure):
    _fields_ = [("act_d", c_void_p), ("act_p", c_void_p), ("noise", c_void_p), ("seed", c_uint64),
                ("auto_reset", c_int32), ("flags", c_int32),
                ("step_count", c_void_p),
                ("reward", c_void_p), ("r_d", c_void_p), ("r_p", c_void_p), ("r_j", c_void_p),
                ("reward64", c_void_p), ("terminated", c_void_p),
                ("pd", c_void_p), ("detected", c_void_p), ("tracking", c_void_p),
                ("snr0", c_void_p), ("snr1", c_void_p), ("jsr_db", c_void_p), ("pd_net", c_void_p),
                ("jam_power", c_void_p),
                ("state", c_void_p), ("obs", c_void_p), ("avail", c_void_p), ("env_begin", c_int32), ("env_count", c_int32)]


class AgentWeights(C.Structure):
    _fields_ = [("obs_dim", c_int32), ("obs_pad", c_int32), ("hidden", c_int32), ("actor_hidden", c_int32),
                ("n_actions", c_int32), ("tc_format", c_int32)] + [
        (n, c_void_p) for n in ("wa1t", "ba1", "wa2t", "ba2", "wa3t", "ba3", "wfc1t", "bfc1", "wrzt", "brz",
                                "wint", "bin", "whnt", "bhn", "wqt", "bq1", "w1a", "w1p", "w2", "bq2", "tc_chunks", "wiht", "whht",
                                "rec_chunks", "bgx")]


class AgentIO(C.Structure):
    _fields_ = [("n_rows", c_int32), ("n_steps", c_int32), ("obs", c_void_p), ("hidden", c_void_p),
                ("hidden_zero_init", c_int32), ("test_mode", c_int32), ("tile_rows", c_int32), ("path", c_int32),
                ("hidden_seq", c_void_p), ("q_all", c_void_p), ("params_all", c_void_p), ("greedy", c_void_p),
                ("sel_actions", c_void_p), ("q_sel", c_void_p), ("avail", c_void_p), ("u_eps", c_void_p),
                ("rand_actions", c_void_p), ("epsilon", c_float), ("rng_step", C.c_uint32), ("seed", c_uint64),
                ("actions", c_void_p), ("power", c_void_p), ("q_chosen", c_void_p), ("part", c_int32), ("rng_row_offset", c_int32),
                ("gate_x", c_void_p), ("epsilon_dev", c_void_p), ("rng_step_dev", c_void_p),
                ("actions_mirror", c_void_p), ("power_mirror", c_void_p), ("hidden_in", c_void_p),
                ("obs_group", c_int32), ("reserved2", c_int32)]


MAX_COPY_KEYS = 12    # include/macjd.h: MACJD_MAX_COPY_KEYS
HOST_PINNED = 1       # include/macjd.h: MACJD_HOST_PINNED
ENV_FOLLOWS_AGENT = 1 # include/macjd.h: MACJD_ENV_FOLLOWS_AGENT


class ActHost(C.Structure):
    _fields_ = [("obs", c_void_p), ("avail", c_void_p), ("actions", c_void_p), ("power", c_void_p), ("q_chosen", c_void_p),
                ("flags", C.c_uint32), ("reserved", C.c_uint32)]


class EnvHost(C.Structure):
    _fields_ = [("act_d", c_void_p), ("act_p", c_void_p), ("reward", c_void_p), ("terminated", c_void_p),
                ("obs", c_void_p), ("state", c_void_p), ("flags", C.c_uint32), ("reserved", C.c_uint32)]


class CopyDesc(C.Structure):
    _fields_ = [("src", c_void_p), ("dst", c_void_p), ("src_ep_stride", c_int64), ("src_t_stride", c_int64),
                ("dst_ep_stride", c_int64), ("dst_t_stride", c_int64), ("n_t", c_int32), ("inner_bytes", c_int32),
                ("vec_bytes", c_int32), ("convert", c_int32)]


class MixerDims(C.Structure):
    _fields_ = [("n_rows", c_int32), ("state_dim", c_int32), ("n_agents", c_int32), ("embed_dim", c_int32),
                ("hyper_hidden", c_int32), ("reserved", c_int32)]


MIXER_FIELDS = ("ln_w", "ln_b", "w1a_w", "w1a_b", "w1b_w", "w1b_b", "wfa_w", "wfa_b", "wfb_w", "wfb_b",
                "b1_w", "b1_b", "va_w", "va_b", "vb_w", "vb_b")
# state_dict key of every macjd_mixer_params field (core/networks.py QMixer)
MIXER_KEYS = ("state_norm.weight", "state_norm.bias", "hyper_w_1.0.weight", "hyper_w_1.0.bias",
              "hyper_w_1.2.weight", "hyper_w_1.2.bias", "hyper_w_final.0.weight", "hyper_w_final.0.bias",
              "hyper_w_final.2.weight", "hyper_w_final.2.bias", "hyper_b_1.weight", "hyper_b_1.bias",
              "V.0.weight", "V.0.bias", "V.2.weight", "V.2.bias")


class MixerParams(C.Structure):
    _fields_ = [(n, c_void_p) for n in MIXER_FIELDS]


class QheadDims(C.Structure):
    _fields_ = [("n_rows", c_int32), ("hidden", c_int32), ("n_actions", c_int32), ("reserved", c_int32)]


MAX_OPT_TENSORS = 32


class OptTensors(C.Structure):
    _fields_ = [("count", c_int32), ("reserved", c_int32), ("param", c_void_p * MAX_OPT_TENSORS),
                ("numel", c_int64 * MAX_OPT_TENSORS)]


# order must match macjd_abi_sizeof() in csrc/macjd_api.cu
ABI_STRUCTS = [Ctx, EnvTables, EnvIO, AgentWeights, AgentIO, CopyDesc, MixerDims, MixerParams, QheadDims, OptTensors,
               ActHost, EnvHost]


class MacjdError(RuntimeError):
    pass


def ptr(x):
    """Device (or, for the host-emulation test library, host) address of a buffer."""
    if x is None:
        return None
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "ctypes"):
        return x.ctypes.data
    return int(x)


class NativeLib:
    """A loaded macjd C-ABI library."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise MacjdError(
                f"macjd CUDA library not found at {path}; build it first "
                f"(python __graft_entry__.py). There is no CPU fallback.")
        self.path = path
        self.lib = C.CDLL(path)
        L = self.lib
        L.macjd_status_string.restype = C.c_char_p
        L.macjd_status_string.argtypes = [C.c_int]
        L.macjd_last_cuda_error.restype = C.c_char_p
        L.macjd_abi_version.restype = C.c_int
        L.macjd_abi_sizeof.restype = C.c_size_t
        L.macjd_abi_sizeof.argtypes = [C.c_int]
        for name, args in self.SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = C.c_int
            fn.argtypes = [C.POINTER(a) for a in args]
        P = C.POINTER
        vp, sz, i32, i64, f32 = c_void_p, C.c_size_t, c_int32, c_int64, c_float
        for name, restype, argtypes in (
            ("macjd_replay_copy", C.c_int, [P(Ctx), P(CopyDesc), i32, vp, i32, i32]),
            ("macjd_env_derived_bytes", sz, [P(EnvTables)]),
            ("macjd_gemm", C.c_int, [P(Ctx), i32, i32, i32, vp, i32, i32, vp, i32, i32, vp, i32, vp, i32, i32, vp, sz]),
            ("macjd_agent_unroll_workspace_floats", sz, [P(AgentWeights), i32, i32]),
            ("macjd_agent_unroll", C.c_int, [P(Ctx), P(AgentWeights), P(AgentIO), vp, sz]),
            ("macjd_env_prepare", C.c_int, [P(Ctx), P(EnvTables), vp]),
            ("macjd_mixer_workspace_floats", sz, [P(MixerDims)]),
            ("macjd_mixer_forward", C.c_int, [P(Ctx), P(MixerDims), P(MixerParams), vp, vp, vp, vp, sz]),
            ("macjd_mixer_backward", C.c_int, [P(Ctx), P(MixerDims), P(MixerParams), vp, vp, vp, sz, P(MixerParams), vp]),
            ("macjd_qhead_scratch_floats", sz, [P(QheadDims)]),
            ("macjd_qhead_forward", C.c_int, [P(Ctx), P(QheadDims), P(AgentWeights), vp, vp, vp, vp, vp]),
            ("macjd_qhead_forward_ws", C.c_int, [P(Ctx), P(QheadDims), P(AgentWeights), vp, vp, vp, vp, vp, vp, C.c_size_t]),
            ("macjd_qhead_backward", C.c_int, [P(Ctx), P(QheadDims), P(AgentWeights), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz]),
            ("macjd_qhead_repack", C.c_int, [P(Ctx), P(AgentWeights), vp, vp, vp, vp, vp, i32, vp, vp, i32]),
            ("macjd_gather_q", C.c_int, [P(Ctx), i32, i32, vp, vp, vp]),
            ("macjd_td_scratch_floats", sz, [i32]),
            ("macjd_td_loss", C.c_int, [P(Ctx), i32, vp, vp, vp, vp, vp, f32, vp, vp, vp, vp, sz]),
            ("macjd_opt_scratch_floats", sz, []),
            ("macjd_tc_gemm_selftest", C.c_int, [P(Ctx), i32, i32, i32, vp, vp, vp]),
            ("macjd_clip_adam", C.c_int, [P(Ctx), P(OptTensors), vp, vp, vp, vp, f32, f32, f32, f32, f32, i64, vp, vp, sz]),
            ("macjd_clip_adam_dev", C.c_int, [P(Ctx), P(OptTensors), vp, vp, vp, vp, f32, f32, f32, f32, f32, vp, vp, vp, sz]),
        ):
            fn = getattr(L, name)
            fn.restype = restype
            fn.argtypes = argtypes
        self._check_abi()

    SIGNATURES = {
        "macjd_env_step": (Ctx, EnvTables, EnvIO),
        "macjd_env_reset": (Ctx, EnvTables, EnvIO),
        "macjd_agent_forward": (Ctx, AgentWeights, AgentIO),
        "macjd_rollout_step": (Ctx, AgentWeights, AgentIO, EnvTables, EnvIO),
        "macjd_rollout_steps": (Ctx, AgentWeights, AgentIO, EnvTables, EnvIO),
        "macjd_agent_act_host": (Ctx, AgentWeights, AgentIO, ActHost),
        "macjd_env_step_host": (Ctx, EnvTables, EnvIO, EnvHost),
        "macjd_rollout_step_host": (Ctx, AgentWeights, AgentIO, ActHost, EnvTables, EnvIO, EnvHost),
    }

    def _check_abi(self):
        if self.lib.macjd_abi_version() != 2:
            raise MacjdError("macjd ABI version mismatch")
        for i, st in enumerate(ABI_STRUCTS):
            got = self.lib.macjd_abi_sizeof(i)
            if got != C.sizeof(st):
                raise MacjdError(f"ABI struct {st.__name__}: library says {got} bytes, binding has {C.sizeof(st)}")

    def __deepcopy__(self, memo):
        return self          # a loaded library is shared, never copied (target networks deepcopy the MAC)

    def __reduce__(self):
        return (NativeLib, (self.path,))

    def check(self, status):
        if status != 0:
            msg = self.lib.macjd_status_string(status).decode()
            if status == -3:
                msg += ": " + self.lib.macjd_last_cuda_error().decode()
            raise MacjdError(f"macjd call failed ({status}): {msg}")

    def call(self, name, *structs):
        # argtypes are POINTER(struct): ctypes passes the instances by reference itself
        status = getattr(self.lib, name)(*structs)
        if status != 0:
            self.check(status)

    def callv(self, name, *args):
        """Entry points with scalar / raw-pointer arguments: structures are passed by
        reference, tensors and arrays by address."""
        conv = []
        for a in args:
            if isinstance(a, C.Structure):
                conv.append(C.byref(a))
            elif isinstance(a, C.Array):
                conv.append(a)
            elif a is None or isinstance(a, (int, float)):
                conv.append(a)
            else:
                conv.append(ptr(a))
        self.check(getattr(self.lib, name)(*conv))


_lib = None
_lock = threading.Lock()


def get_lib() -> NativeLib:
    """The product library (built in-tree). Raises if it is missing."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                _lib = NativeLib(LIB_PATH)
    return _lib


def torch_ctx(device=None):
    """macjd_ctx for torch's current stream on ``device``."""
    import torch
    idx = getattr(device, "index", None) if device is not None and not isinstance(device, (str, int)) else None
    if idx is None:
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
    raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)      # same handle, without building a Stream object
    stream = raw(idx) if raw is not None else torch.cuda.current_stream(idx).cuda_stream
    c = _ctx_cache.get((idx, stream))                                # one struct per (device, stream): ~1 us per act call
    if c is None:
        c = _ctx_cache[(idx, stream)] = Ctx(device=idx, reserved=0, stream=stream)
    return c


_ctx_cache = {}
