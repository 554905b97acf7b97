"""Shared helpers for the parity tests (test infrastructure)."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_env_golden(name):
    g = np.load(os.path.join(GOLDEN, f"env_{name}.npz"))
    cfg = json.loads(str(g["config_json"]))
    return g, cfg


def spec_for_golden(g, cfg, n_envs=1):
    from macjd_b200.simulation.scenario import spec_from_config
    return spec_from_config(cfg, n_envs=n_envs, episode_limit=int(g["episode_limit"]))


_EMUL = None


def emul_lib():
    """The kernel sources compiled for the host (tests/emul): same C ABI, host memory."""
    global _EMUL
    if _EMUL is None:
        from tests.emul.build_emul import build
        from macjd_b200 import _native as N
        _EMUL = N.NativeLib(build())
    return _EMUL


def has_cuda():
    import torch
    return torch.cuda.is_available()
