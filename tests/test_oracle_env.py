"""The NumPy environment oracle against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import json

import numpy as np
import pytest

from oracle.env_oracle import (EnvOracle, albersheim_pd, compact_noise_for_reference,
                               device_noise, philox4x32_10)
from tests.helpers import load_env_golden, spec_for_golden

SCEN = ["default", "selftest", "active", "rand0", "rand1", "rand2", "rand3"]      # rand*: random scenarios (make_golden.py env_random)
RTOL = 1e-12  # float64 restatement vs float64 reference: operation-order noise only


@pytest.mark.parametrize("name", SCEN)
def test_static_views_match_reference(name):
    g, cfg = load_env_golden(name)
    env = EnvOracle(spec_for_golden(g, cfg))
    assert env.get_env_info() == json.loads(str(g["env_info_json"]))
    s = env.reset()
    assert s.dtype == np.float32 and s.shape == (1, env.state_dim)
    np.testing.assert_array_equal(s[0], g["state0"][0])
    np.testing.assert_array_equal(env.get_obs()[0], g["obs0"][0])
    av = env.get_avail_actions()
    assert av.dtype == np.int32
    np.testing.assert_array_equal(av[0], g["avail0"][0])


@pytest.mark.parametrize("name", SCEN)
def test_entity_constants(name):
    g, cfg = load_env_golden(name)
    env = EnvOracle(spec_for_golden(g, cfg))
    np.testing.assert_array_equal(env.gt[0], g["radar_gt_lin"])
    np.testing.assert_array_equal(env.pn[0], g["radar_pn_watts"])
    np.testing.assert_array_equal(env.gj[0], g["jammer_gj_lin"])
    np.testing.assert_allclose(env.echo_power()[0, :, 0], g["echo_ps"], rtol=RTOL)
    np.testing.assert_allclose(albersheim_pd(g["pd_table_snr"]), g["pd_table"], rtol=RTOL)


@pytest.mark.parametrize("name", SCEN)
def test_step_sequences_match_reference(name):
    g, cfg = load_env_golden(name)
    check_step_sequences(g, cfg)


def check_step_sequences(g, cfg):
    """The oracle against one recording of the reference (the arrays of tests/golden/make_golden.py:record_env)."""
    n_eps, T = g["reward"].shape
    # all golden episodes advance side by side as a batch of n_eps envs
    env = EnvOracle(spec_for_golden(g, cfg, n_envs=n_eps))
    env.reset()
    for t in range(T):
        out = env.step(g["act_d"][:, t], g["act_p"][:, t], g["noise"][:, t])
        for key in ("reward", "r_d", "r_p", "r_j"):
            np.testing.assert_allclose(out[key], g[key][:, t], rtol=1e-10, atol=1e-13, err_msg=f"{key} t={t}")
        np.testing.assert_allclose(out["pd"][:, :, 0], g["pd"][:, t], rtol=RTOL)
        np.testing.assert_allclose(out["snr0"][:, :, 0], g["snr0"][:, t], rtol=RTOL)
        np.testing.assert_allclose(out["snr1"][:, :, 0], g["snr1"][:, t], rtol=RTOL)
        np.testing.assert_array_equal(out["tracking"], g["tracking"][:, t])
        np.testing.assert_array_equal(out["terminated"], g["terminated"][:, t])
        np.testing.assert_array_equal(out["step_count"], g["step_count"][:, t])
        # RNG consumption of the reference: R draws + one per valid deception action
        np.testing.assert_array_equal(env.R + out["deception"].sum(axis=1), g["n_rng_draws"][:, t])


def test_survey_known_answers():
    """Values observed on the reference in SURVEY.md section 8c (default scenario)."""
    g, cfg = load_env_golden("default")
    env = EnvOracle(spec_for_golden(g, cfg))
    np.testing.assert_allclose(env.echo_power()[0, :, 0], [3.3534683110189373e-10, 2.0120809866113628e-11], rtol=1e-13)
    np.testing.assert_allclose(env.pn[0], 0.001995262314968879, rtol=1e-15)
    np.testing.assert_allclose(albersheim_pd([0, 0.5, 1, 2, 5, 80]),
                               [0.10292481173711425, 0.34294051008939636, 0.7036428926615178,
                                0.980053481878016, 0.9999977034328043, 1.0], rtol=1e-13)
    env.reset()
    o = env.step([[0, 0]], [[0.3, 0.9]], [[0.05, 0.5, 0.0, 0.0]])
    np.testing.assert_allclose([o["reward"][0], o["r_d"][0], o["r_p"][0], o["r_j"][0]], [-0.928, -0.8, -0.128, 0.0], rtol=1e-12, atol=1e-15)
    assert o["tracking"][0].tolist() == [True, False]
    o = env.step([[4, 4]], [[0.3, 0.9]], [[0.2, 0.05, 0.5, 0.5]])
    np.testing.assert_allclose([o["r_d"][0], o["r_p"][0], o["r_j"][0], o["reward"][0]],
                               [-1.2, -0.128, 0.999999999999, -0.3280000000009998], rtol=1e-11)
    o = env.step([[1, 2]], [[0.0, 1e-9]], [[0.2, 0.2, 0.9, 0.0]])
    np.testing.assert_allclose([o["r_p"][0], o["r_j"][0], o["reward"][0]],
                               [-0.02000000009, 0.10292617637767454, 0.08292617628767454], rtol=1e-11)
    np.testing.assert_allclose(o["prj"][0, 1], 9.732986902287215e-10, rtol=1e-13)
    o = env.step([[7, 0]], [[0.5, 0.0]], [[0.9, 0.9, 0.0, 0.0]])
    np.testing.assert_allclose(o["reward"][0], -0.065, rtol=1e-12)


def test_multi_target_reduces_to_single_target():
    """K > 1 extension: with identical targets every per-pair quantity is replicated and
    the K == 1 outputs are recovered when the extra targets' draws never detect."""
    g, cfg = load_env_golden("active")
    spec1 = spec_for_golden(g, cfg, n_envs=3)
    specK = spec_for_golden(g, cfg, n_envs=3)
    K = 3
    specK["targets"]["position"] = np.repeat(specK["targets"]["position"], K, axis=1)
    specK["targets"]["rcs"] = np.repeat(specK["targets"]["rcs"], K, axis=1)
    e1, eK = EnvOracle(spec1), EnvOracle(specK)
    rng = np.random.default_rng(5)
    R, J = e1.R, e1.J
    for t in range(20):
        act_d = rng.integers(0, 2 * R + 1, size=(3, J))
        act_p = rng.random((3, J))
        n1 = rng.random((3, R + J))
        nK = np.ones((3, R * K + J)) * 2.0          # u = 2 never detects
        nK[:, 0:R * K:K] = n1[:, :R]
        nK[:, R * K:] = n1[:, R:]
        o1, oK = e1.step(act_d, act_p, n1), eK.step(act_d, act_p, nK)
        np.testing.assert_allclose(oK["pd"][:, :, 0], o1["pd"][:, :, 0], rtol=1e-15)
        np.testing.assert_array_equal(oK["tracking"], o1["tracking"])
        np.testing.assert_allclose(oK["r_d"], o1["r_d"])
        np.testing.assert_allclose(oK["r_p"], o1["r_p"])
        # suppression reward sums the per-pair reductions: K identical targets -> K x
        assert np.all(oK["r_j"] >= o1["r_j"] - 1e-12)
        np.testing.assert_allclose(oK["pd_net"][:, 0], 1 - np.prod(1 - o1["pd"][:, :, 0], axis=1))


def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10."""
    assert philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert philox4x32_10((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    assert philox4x32_10((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == \
        (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)
    u = device_noise(7, 1, 3, 5, 6)
    assert u.dtype == np.float32 and np.all((u >= 0) & (u < 1))


def test_compact_noise_helper():
    seq = compact_noise_for_reference(np.array([.1, .2, .3, .4, .5]), [False, True, True], R=2)
    assert seq == [.1, .2, .4, .5]
