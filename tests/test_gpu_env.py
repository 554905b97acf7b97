"""GPU parity tests of the fused environment step kernel, through the C ABI
(libmacjd_b200.so) and the drop-in Python class."""
import types

import numpy as np
import pytest
import torch

from tests.env_checks import check_env_against_golden, check_env_against_oracle, check_shim_types

pytestmark = pytest.mark.gpu


def make_env(spec, **kw):
    from macjd_b200.simulation.environment import ElectromagneticEnvironment
    return ElectromagneticEnvironment(types.SimpleNamespace(), spec=spec, device="cuda", **kw)


@pytest.mark.parametrize("name", ["default", "selftest", "active", "rand0", "rand1", "rand2", "rand3"])
def test_env_kernel_vs_reference_golden(name):
    check_env_against_golden(make_env, name)


@pytest.mark.parametrize("kind,n", [("hetero", 4096), ("active", 5000), ("scaled", 1111), ("scaled_small", 777),
                                    ("odd_radars", 1000), ("one_radar", 333), ("active", 20000)])
def test_env_kernel_vs_oracle(kind, n):
    check_env_against_oracle(make_env, kind, n, steps=5)


def test_shim_returns_reference_types():
    from macjd_b200.simulation.environment import ElectromagneticEnvironment
    check_shim_types(lambda path, rl: ElectromagneticEnvironment(rl, sim_config_path=path))


def test_full_size_properties():
    """BASELINE config 2/3 sizes: size-independent properties instead of the oracle --
    replicated scenarios give identical rows for identical inputs; idle actions at zero
    power give r_j = 0, r_p = J * rp_max; obs == state replicated; termination at the limit."""
    from macjd_b200.simulation.scenario import default_spec, scaled_spec
    for spec, n in ((default_spec(1 << 18, episode_limit=3), 1 << 18), (scaled_spec(8192, episode_limit=3, seed=1), 8192)):
        env = make_env(spec)
        J, R, K = env.num_jammers, env.num_radars, env.num_targets
        z = torch.zeros(n, J, dtype=torch.int32, device="cuda")
        p = torch.zeros(n, J, device="cuda")
        noise = torch.full((n, R * K + J), 2.0, device="cuda")        # u = 2: nothing is ever detected
        for t in range(1, 4):
            obs, reward, term, info = env.step((z, p), noise=noise)
            assert bool((term == (t >= 3)).all())
            assert bool((info["r_j"] == 0).all()) and bool((info["r_d"] == 0).all())
            torch.testing.assert_close(info["r_p"], torch.full((n,), J * -0.01, device="cuda"), rtol=1e-6, atol=0)
            assert bool((obs == env.state[:, None, :]).all())
            assert bool((env.avail == 1).all())
        noise.zero_()                                                 # u = 0: everything is detected
        _, _, _, info = env.step((z, p), noise=noise)
        assert bool(info["radar_tracking"].all())


@pytest.mark.parametrize("kind,n", [("active", 5000), ("scaled", 1111), ("scaled_small", 777), ("odd_radars", 1000),
                                    ("one_radar", 333), ("hetero", 40000)])
@pytest.mark.parametrize("uniform_min", [1, 1 << 30])
def test_raw_and_derived_table_kernels_agree(kind, n, uniform_min, monkeypatch):
    """The step kernel on derived scenario tables (csrc/env_step2.cuh) against the raw-table kernel (csrc/env_step.cuh)."""
    # both block forms of the derived-table kernel: split roles (small batches) and uniform roles (>= 16 384 envs)
    monkeypatch.setenv("MACJD_ENV_UNIFORM_MIN", str(uniform_min))
    from tests.env_checks import check_raw_and_derived_kernels_agree
    check_raw_and_derived_kernels_agree(make_env, kind, n)
