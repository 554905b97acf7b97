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
LIB_PATH = os.path.join(_HERE, "csrc", "libmacjd_b200.so")

c_void_p, c_int32, c_int64, c_uint64, c_double, c_float = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double, C.c_float


class Ctx(C.Structure):
    _fields_ = [("device", c_int32), ("reserved", c_int32), ("stream", c_void_p)]


class EnvTables(C.Structure):
    _fields_ = [("n_envs", c_int32), ("n_jammers", c_int32), ("n_radars", c_int32), ("n_targets", c_int32),
                ("n_types", c_int32), ("episode_limit", c_int32),
                ("data", c_void_p), ("row_stride", c_int64), ("env_stride", c_int64),
                ("rd_min", c_double), ("rd_max", c_double), ("rp_min", c_double), ("rp_max", c_double),
                ("alb_a", c_double), ("alb_zoff", c_double), ("alb_den", c_double)]


class EnvIO(C.Structure):
    _fields_ = [("act_d", c_void_p), ("act_p", c_void_p), ("noise", c_void_p), ("seed", c_uint64),
                ("auto_reset", c_int32), ("reserved", c_int32),
                ("step_count", c_void_p),
                ("reward", c_void_p), ("r_d", c_void_p), ("r_p", c_void_p), ("r_j", c_void_p),
                ("reward64", c_void_p), ("terminated", c_void_p),
                ("pd", c_void_p), ("detected", c_void_p), ("tracking", c_void_p),
                ("snr0", c_void_p), ("snr1", c_void_p), ("jsr_db", c_void_p), ("pd_net", c_void_p),
                ("jam_power", c_void_p),
                ("state", c_void_p), ("obs", c_void_p), ("avail", c_void_p)]


class AgentWeights(C.Structure):
    _fields_ = [("obs_dim", c_int32), ("obs_pad", c_int32), ("hidden", c_int32), ("actor_hidden", c_int32),
                ("n_actions", c_int32), ("reserved", c_int32)] + [
        (n, c_void_p) for n in ("wa1t", "ba1", "wa2t", "ba2", "wa3t", "ba3", "wfc1t", "bfc1", "wrzt", "brz",
                                "wint", "bin", "whnt", "bhn", "wqt", "bq1", "w1a", "w1p", "w2", "bq2")]


class AgentIO(C.Structure):
    _fields_ = [("n_rows", c_int32), ("n_steps", c_int32), ("obs", c_void_p), ("hidden", c_void_p),
                ("hidden_zero_init", c_int32), ("test_mode", c_int32), ("tile_rows", c_int32), ("reserved", c_int32),
                ("hidden_seq", c_void_p), ("q_all", c_void_p), ("params_all", c_void_p), ("greedy", c_void_p),
                ("sel_actions", c_void_p), ("q_sel", c_void_p), ("avail", c_void_p), ("u_eps", c_void_p),
                ("rand_actions", c_void_p), ("epsilon", c_float), ("rng_step", C.c_uint32), ("seed", c_uint64),
                ("actions", c_void_p), ("power", c_void_p), ("q_chosen", c_void_p)]


# order must match macjd_abi_sizeof() in csrc/macjd_api.cu
ABI_STRUCTS = [Ctx, EnvTables, EnvIO, AgentWeights, AgentIO]


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
        self._check_abi()

    SIGNATURES = {
        "macjd_env_step": (Ctx, EnvTables, EnvIO),
        "macjd_env_reset": (Ctx, EnvTables, EnvIO),
        "macjd_agent_forward": (Ctx, AgentWeights, AgentIO),
    }

    def _check_abi(self):
        if self.lib.macjd_abi_version() != 1:
            raise MacjdError("macjd ABI version mismatch")
        for i, st in enumerate(ABI_STRUCTS):
            got = self.lib.macjd_abi_sizeof(i)
            if got != C.sizeof(st):
                raise MacjdError(f"ABI struct {st.__name__}: library says {got} bytes, binding has {C.sizeof(st)}")

    def check(self, status):
        if status != 0:
            msg = self.lib.macjd_status_string(status).decode()
            if status == -3:
                msg += ": " + self.lib.macjd_last_cuda_error().decode()
            raise MacjdError(f"macjd call failed ({status}): {msg}")

    def call(self, name, *structs):
        self.check(getattr(self.lib, name)(*[C.byref(s) for s in structs]))


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
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    return Ctx(device=idx, reserved=0, stream=torch.cuda.current_stream(idx).cuda_stream)
