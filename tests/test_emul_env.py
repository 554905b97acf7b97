"""Kernel logic of csrc/env_step.cuh checked on the CPU: the same CUDA source compiled
for the host (tests/emul) behind the same C ABI and the same Python class, against the
reference goldens and the NumPy oracle.  The GPU run of the same checks is
tests/test_gpu_env.py."""
import types

import numpy as np
import pytest
import torch

from oracle.env_oracle import EnvOracle, device_noise
from tests.helpers import emul_lib, load_env_golden, spec_for_golden
from tests.env_checks import check_env_against_golden, check_env_against_oracle, check_shim_types

SCEN = ["default", "selftest", "active", "rand0", "rand1", "rand2", "rand3"]      # rand*: random scenarios (make_golden.py env_random)


def make_env(spec, **kw):
    from macjd_b200.simulation.environment import ElectromagneticEnvironment
    return ElectromagneticEnvironment(types.SimpleNamespace(), spec=spec, device="cpu", _lib=emul_lib(), **kw)


@pytest.mark.parametrize("name", SCEN)
def test_env_kernel_vs_reference_golden(name):
    check_env_against_golden(make_env, name)


@pytest.mark.parametrize("kind,n", [("hetero", 200), ("active", 133), ("scaled", 70)])
def test_env_kernel_vs_oracle(kind, n):
    check_env_against_oracle(make_env, kind, n, steps=6)


def test_shared_scenario_tables_equal_per_env_tables():
    from macjd_b200.simulation.scenario import default_spec
    spec = default_spec(65)
    a, b = make_env(spec), make_env(spec, share_scenario=True)
    rng = np.random.default_rng(0)
    for _ in range(3):
        act_d = torch.from_numpy(rng.integers(0, 5, size=(65, 2)).astype(np.int32))
        act_p = torch.from_numpy(rng.random((65, 2)).astype(np.float32))
        noise = torch.from_numpy(rng.random((65, 4)).astype(np.float32))
        a.step((act_d, act_p), noise=noise)
        b.step((act_d, act_p), noise=noise)
        for k in ("reward", "pd", "detected", "state", "obs", "avail", "terminated"):
            assert torch.equal(getattr(a, k), getattr(b, k)), k


def test_device_philox_noise_matches_oracle_mapping():
    from macjd_b200.simulation.scenario import hetero_spec
    spec = hetero_spec(40, seed=3, active=True)
    env = make_env(spec, seed=0xDEADBEEFCAFE)
    ora = EnvOracle(spec)
    rng = np.random.default_rng(1)
    R, J, K = ora.R, ora.J, ora.K
    for t in range(1, 4):
        act_d = rng.integers(0, 2 * R + 1, size=(40, J)).astype(np.int32)
        act_p = rng.random((40, J)).astype(np.float32)
        noise = np.stack([device_noise(0xDEADBEEFCAFE, 1, e, t, R * K + J) for e in range(40)])
        o = ora.step(act_d, act_p, noise)
        env.step((torch.from_numpy(act_d), torch.from_numpy(act_p)))   # no injected noise -> Philox
        np.testing.assert_array_equal(env.detected.numpy().T.reshape(40, R, K).astype(bool), o["detected"])
        np.testing.assert_allclose(env.reward64.numpy(), o["reward"], rtol=1e-12, atol=1e-14)


def test_shim_returns_reference_types():
    check_shim_types(lambda cfg_path, rl: __import__("macjd_b200.simulation.environment", fromlist=["x"])
                     .ElectromagneticEnvironment(rl, sim_config_path=cfg_path, device="cpu", _lib=emul_lib()))


def test_auto_reset_and_termination():
    from macjd_b200.simulation.scenario import default_spec
    env = make_env(default_spec(3, episode_limit=3), auto_reset=True)
    z = torch.zeros(3, 2, dtype=torch.int32)
    p = torch.zeros(3, 2)
    for t in range(1, 8):
        _, _, term, info = env.step((z, p))
        assert term.tolist() == [t % 3 == 0] * 3
        assert info["step_count"].tolist() == [t % 3] * 3


def test_invalid_arguments_are_status_codes():
    from macjd_b200 import _native as N
    lib = emul_lib()
    with pytest.raises(N.MacjdError, match="invalid argument"):
        lib.call("macjd_env_step", N.Ctx(), N.EnvTables(n_envs=4, n_jammers=2, n_radars=2, n_targets=1, n_types=4), N.EnvIO())
    with pytest.raises(N.MacjdError, match="unsupported"):
        lib.call("macjd_env_step", N.Ctx(), N.EnvTables(n_envs=4, n_jammers=2, n_radars=65, n_targets=1, n_types=4), N.EnvIO())


@pytest.mark.parametrize("kind,n", [("active", 70), ("scaled_small", 45), ("odd_radars", 33)])
@pytest.mark.parametrize("uniform_min", [1, 1 << 30])
def test_raw_and_derived_table_kernels_agree(kind, n, uniform_min, monkeypatch):
    # both block forms of the derived-table kernel: split roles (small batches) and uniform roles (>= 16 384 envs)
    monkeypatch.setenv("MACJD_ENV_UNIFORM_MIN", str(uniform_min))
    from tests.env_checks import check_raw_and_derived_kernels_agree
    check_raw_and_derived_kernels_agree(make_env, kind, n)
