"""Environment parity checks shared by the host-emulation (CPU) and the GPU test files."""
import json
import os
import tempfile
import types

import numpy as np
import torch
import yaml

from oracle.env_oracle import EnvOracle
from tests.helpers import load_env_golden, spec_for_golden

F32_RTOL = 2e-7   # outputs are float64 results rounded to float32 (half an ulp = 6e-8)


def _np(t):
    return t.detach().cpu().numpy()


def check_env_against_golden(make_env, name):
    """Drive the kernel with the exact action / noise sequences the unmodified reference
    saw: integer outputs bit-exact, float64 reward to 1e-10, float32 views to an ulp."""
    g, cfg = load_env_golden(name) if isinstance(name, str) else name        # (or a recording: (arrays, config))
    n_eps, T = g["reward"].shape
    env = make_env(spec_for_golden(g, cfg, n_envs=n_eps))
    dev = env.device
    info = env.get_env_info()
    assert info == json.loads(str(g["env_info_json"]))
    s0 = env.reset()
    np.testing.assert_array_equal(_np(s0), g["state0"])
    np.testing.assert_array_equal(_np(env.get_obs()), g["obs0"])
    np.testing.assert_array_equal(_np(env.get_avail_actions()), g["avail0"].astype(np.uint8))
    R = env.num_radars
    for t in range(T):
        act = (torch.from_numpy(g["act_d"][:, t]).to(dev), torch.from_numpy(g["act_p"][:, t]).to(dev))
        obs, reward, term, inf = env.step(act, noise=torch.from_numpy(g["noise"][:, t]).to(dev))
        np.testing.assert_array_equal(_np(inf["radar_tracking"]), g["tracking"][:, t], err_msg=f"tracking t={t}")
        np.testing.assert_array_equal(_np(term), g["terminated"][:, t])
        np.testing.assert_array_equal(_np(inf["step_count"]), g["step_count"][:, t])
        np.testing.assert_allclose(_np(env.reward64), g["reward"][:, t], rtol=1e-10, atol=1e-13, err_msg=f"reward t={t}")
        for key in ("r_d", "r_p", "r_j"):
            np.testing.assert_allclose(_np(inf[key]), g[key][:, t], rtol=1e-6, atol=1e-7, err_msg=f"{key} t={t}")
        np.testing.assert_allclose(_np(reward), g["reward"][:, t], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(_np(inf["radar_pds"])[:, :, 0], g["pd"][:, t], rtol=F32_RTOL)
        np.testing.assert_allclose(_np(inf["snr_no_jamming"])[:, :, 0], g["snr0"][:, t], rtol=F32_RTOL)
        np.testing.assert_allclose(_np(inf["snr_with_jamming"])[:, :, 0], g["snr1"][:, t], rtol=F32_RTOL)
        np.testing.assert_array_equal(_np(obs), g["obs0"])           # observations are static


def make_spec(kind, n):
    from macjd_b200.simulation.scenario import hetero_spec, scaled_spec
    if kind == "hetero":
        return hetero_spec(n, seed=21)
    if kind == "active":
        return hetero_spec(n, seed=22, active=True)
    if kind == "scaled":
        return scaled_spec(n, seed=23)
    if kind == "scaled_small":
        return scaled_spec(n, n_jammers=3, n_radars=5, n_targets=2, seed=24)
    if kind == "odd_radars":            # small-batch kernel with two physics workers, radars 0 and 2 / radar 1; S % 4 != 0
        return scaled_spec(n, n_jammers=2, n_radars=3, n_targets=1, seed=25)
    if kind == "one_radar":             # nothing to share between the physics workers
        return scaled_spec(n, n_jammers=2, n_radars=1, n_targets=2, seed=26)
    raise KeyError(kind)


def check_env_against_oracle(make_env, kind, n, steps=5, seed=5):
    """Synthetic scenarios (incl. K > 1 targets, ragged n): every output against the
    float64 NumPy oracle on the same actions / noise."""
    spec = make_spec(kind, n)
    env, ora = make_env(spec), EnvOracle(spec)
    dev = env.device
    R, J, K = ora.R, ora.J, ora.K
    rng = np.random.default_rng(seed)
    np.testing.assert_array_equal(_np(env.reset()), ora.reset())
    np.testing.assert_array_equal(_np(env.get_obs()), ora.get_obs())
    np.testing.assert_array_equal(_np(env.get_avail_actions()), ora.get_avail_actions().astype(np.uint8))
    for t in range(steps):
        act_d = rng.integers(-1, 2 * R + 3, size=(n, J)).astype(np.int32)
        act_p = (rng.random((n, J)) * 1.2 - 0.1).astype(np.float32)
        noise = rng.random((n, R * K + J)).astype(np.float32)
        o = ora.step(act_d, act_p, noise)
        obs, reward, term, inf = env.step((torch.from_numpy(act_d).to(dev), torch.from_numpy(act_p).to(dev)),
                                          noise=torch.from_numpy(noise).to(dev))
        np.testing.assert_array_equal(_np(inf["detected"]), o["detected"], err_msg=f"detected t={t}")
        np.testing.assert_array_equal(_np(inf["radar_tracking"]), o["tracking"])
        np.testing.assert_array_equal(_np(term), o["terminated"])
        np.testing.assert_allclose(_np(env.reward64), o["reward"], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(_np(inf["radar_pds"]), o["pd"], rtol=F32_RTOL)
        np.testing.assert_allclose(_np(inf["snr_with_jamming"]), o["snr1"], rtol=F32_RTOL)
        np.testing.assert_allclose(_np(inf["snr_no_jamming"]), o["snr0"], rtol=F32_RTOL)
        np.testing.assert_allclose(_np(inf["pd_networked"]), o["pd_net"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(_np(inf["jammer_power"]), ora.jammer_power, rtol=F32_RTOL)
        fin = np.isfinite(o["jsr_db"])
        np.testing.assert_allclose(_np(inf["jsr_db"])[fin], o["jsr_db"][fin], rtol=1e-6, atol=1e-5)
        for key in ("r_d", "r_p", "r_j"):
            np.testing.assert_allclose(_np(inf[key]), o[key], rtol=1e-6, atol=1e-7)
    return env, ora


def check_shim_types(make_single_env):
    """n_envs = 1 shim: exact reference types and the survey's known answers."""
    from macjd_b200.simulation.scenario import default_config_dict
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "simulation_config.yaml")
        with open(path, "w") as f:
            yaml.safe_dump(default_config_dict(), f)
        env = make_single_env(path, types.SimpleNamespace())
        assert env.get_env_info() == {"state_shape": 24, "obs_shape": 24, "n_actions": 5, "n_agents": 2, "episode_limit": 100}
        s = env.reset()
        assert isinstance(s, np.ndarray) and s.dtype == np.float32 and s.shape == (24,)
        np.testing.assert_array_equal(s, np.array([300, 2, 5, 0, 1, 0, 0, 0, 400, 0, 180, 1.8, 4, 0, 0, 1, 0, 180, -400, 0,
                                                   50, 50, -50, -50], dtype=np.float32))
        av = env.get_avail_actions()
        assert isinstance(av, list) and len(av) == 2 and av[0].dtype == np.int32 and av[0].tolist() == [1] * 5
        obs, reward, term, info = env.step([(0, .3), (0, .9)], noise=[.05, .5, 0, 0])
        assert isinstance(obs, list) and len(obs) == 2 and obs[0].dtype == np.float32
        assert isinstance(reward, float) and isinstance(term, bool) and term is False
        np.testing.assert_allclose(reward, -0.928, rtol=1e-7)   # act_p is float32 at the ABI (as in the runner)
        assert set(info) == {"radar_pds", "radar_states", "snr_no_jamming", "snr_with_jamming", "r_d", "r_p", "r_j", "jammer_actions"}
        assert [s_["is_tracking"] for s_ in info["radar_states"]] == [True, False]
        _, reward, _, info = env.step([(4, .3), (4, .9)], noise=[.2, .05, .5, .5])
        ja = info["jammer_actions"]
        assert [a["jammer_idx"] for a in ja] == [0, 1] and all(a["type"] == 0 and a["target_idx"] == 1 for a in ja)
        assert all(set(a) == {"jammer_idx", "target_idx", "type", "power", "received_power"} for a in ja)
        np.testing.assert_allclose(reward, -0.3280000000009998, rtol=1e-7)
        np.testing.assert_allclose(info["r_j"], 0.999999999999, rtol=1e-6)
        _, reward, _, _ = env.step([(1, 0.0), (2, 1e-9)], noise=[.2, .2, .9, 0.0])
        np.testing.assert_allclose(reward, 0.08292617628767454, rtol=1e-7)
        _, reward, _, _ = env.step([(7, .5), (0, 0)], noise=[.9, .9, 0, 0])
        np.testing.assert_allclose(reward, -0.065, rtol=1e-7)
        import pytest
        with pytest.raises(ValueError):
            env.step([(0, 0.0)])
        with pytest.raises(ValueError):
            env.get_agent_obs(5)
        assert np.array_equal(env.get_agent_obs(1), s)
        env.close()


def check_raw_and_derived_kernels_agree(make_env, kind, n, steps=4, seed=8):
    """The step kernel on derived scenario tables (macjd_env_prepare; what the Python environment runs) against the
    kernel that works from the raw tables on every step (`derive_tables=False`, round 1's): every integer output
    bit-identical, float64 rewards to 1e-12, Philox draws included (same counters)."""
    spec = make_spec(kind, n)
    a, b = make_env(spec, derive_tables=True), make_env(spec, derive_tables=False)
    assert a._ctab.derived and not b._ctab.derived
    dev = a.device
    rng = np.random.default_rng(seed)
    R, J = a.num_radars, a.num_jammers
    for env in (a, b):
        env.reset()
    for k in ("state", "obs", "avail"):
        np.testing.assert_array_equal(_np(getattr(a, k)), _np(getattr(b, k)), err_msg=k)
    for t in range(steps):
        act_d = torch.from_numpy(rng.integers(-1, 2 * R + 3, size=(n, J)).astype(np.int32)).to(dev)
        act_p = torch.from_numpy((rng.random((n, J)) * 1.2 - 0.1).astype(np.float32)).to(dev)
        for env in (a, b):
            env.step((act_d, act_p))                     # device Philox noise
        for k in ("detected", "tracking", "terminated", "step_count", "state", "obs", "avail", "pd", "snr0", "snr1", "jam_power"):
            np.testing.assert_array_equal(_np(getattr(a, k)), _np(getattr(b, k)), err_msg=f"{k} t={t}")
        np.testing.assert_allclose(_np(a.reward64), _np(b.reward64), rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(_np(a.pd_net), _np(b.pd_net), rtol=1e-6, atol=1e-7)
        fin = np.isfinite(_np(b.jsr_db))
        np.testing.assert_array_equal(np.isfinite(_np(a.jsr_db)), fin)
        np.testing.assert_allclose(_np(a.jsr_db)[fin], _np(b.jsr_db)[fin], rtol=1e-6)
