#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED reference.

Runs only in the build container, where the reference is mounted read-only at
/root/reference (it does not exist on the GPU box; the tests read the committed
.npz / .json files only).  Usage:

    python tests/golden/make_golden.py [env] [agent] [mixer] [learner] [replay] [api]

Inputs are seeded; noise that the reference would draw from np.random / torch RNG
is injected (FIFO patches) so that both sides see identical sequences.
"""
import contextlib
import inspect
import io
import json
import os
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

import yaml  # noqa: E402


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


# ------------------------------------------------------------------ scenarios
def scenario_default():
    with open(os.path.join(REF, "config", "simulation_config.yaml")) as f:
        return yaml.safe_load(f)


def scenario_selftest():
    """The scenario of the reference's own smoke block (simulation/environment.py:588-611):
    pt ~ 1e6 W, pn -90 dBm, 10 km ranges, a jammer with power_min = 10 W."""
    return {
        "radars": [
            dict(pt=1e6, gt=30, gr=30, wavelength=0.03, rcs=1.0, loss=10, latm=2, pn=-90, type_id=0,
                 position=[10000, 0], theta_m=2.0, theta_a=0.0, t_s=5.0,
                 pulse_compression_gain=100.0, anti_jamming_factor=10.0, threat_level=0.8),
            dict(pt=1.2e6, gt=32, gr=32, wavelength=0.03, rcs=1.5, loss=8, latm=2, pn=-92, type_id=1,
                 position=[-10000, 5000], theta_m=1.8, theta_a=180.0, t_s=4.0,
                 pulse_compression_gain=120.0, anti_jamming_factor=15.0, threat_level=1.2),
        ],
        "jammers": [
            dict(power_max=100, power_min=0, gj=20, loss=5, latm=2, bj=1e6, position=[0, 1000]),
            dict(power_max=120, power_min=10, gj=22, loss=4, latm=2, bj=1.2e6, position=[0, -1000]),
        ],
        "protected_target": dict(position=[0, 0], rcs=2.0),
        "environment_params": dict(max_radar_types=4,
                                   rewards=dict(rd_min=-1.2, rd_max=-0.8, rp_min=-0.1, rp_max=-0.01)),
    }


def scenario_active():
    """Build-authored scenario: 3 radars x 3 jammers, Pd in mid-range, jamming powers
    that move Pd, one jammer co-located with a radar (distance 0 -> action has no
    effect), one jammer with power_min == power_max (zero power range), threat levels
    outside [0.8, 1.2] (clipped), 5 radar types."""
    cfg = {
        "radars": [
            dict(pt=2.0e4, gt=28, gr=28, wavelength=0.05, rcs=1.0, loss=6, latm=1, pn=-60, type_id=4,
                 position=[900.0, 100.0], theta_m=2.5, theta_a=10.0, t_s=3.0,
                 pulse_compression_gain=50.0, anti_jamming_factor=4.0, threat_level=0.5),
            dict(pt=3.5e4, gt=26, gr=27, wavelength=0.04, rcs=1.0, loss=7, latm=1.5, pn=-62, type_id=0,
                 position=[-700.0, 650.0], theta_m=1.5, theta_a=200.0, t_s=6.0,
                 pulse_compression_gain=80.0, anti_jamming_factor=8.0, threat_level=1.5),
            dict(pt=1.2e4, gt=30, gr=29, wavelength=0.03, rcs=1.0, loss=5, latm=1, pn=-58, type_id=2,
                 position=[120.0, -800.0], theta_m=3.0, theta_a=90.0, t_s=2.0,
                 pulse_compression_gain=64.0, anti_jamming_factor=2.0, threat_level=1.0),
        ],
        "jammers": [
            dict(power=5, gj=3, loss=2, latm=1, bj=2.0e6, position=[60.0, 40.0], power_min=0.0, power_max=2.0),
            dict(power=5, gj=4, loss=3, latm=1, bj=1.0e6, position=[120.0, -800.0], power_min=0.5, power_max=1.5),
            dict(power=5, gj=2, loss=2, latm=2, bj=3.0e6, position=[-80.0, -30.0], power_min=1.0, power_max=1.0),
        ],
        "protected_target": dict(position=[15.0, -25.0], rcs=3.0),
        "environment_params": dict(max_radar_types=5,
                                   rewards=dict(rd_min=-1.2, rd_max=-0.8, rp_min=-0.1, rp_max=-0.01)),
    }
    # rescale pt so that the un-jammed SNR Ga*Ps/Pn is 1.5 / 0.8 / 2.5 (Pd 0.91 / 0.56 / 0.99)
    tx, ty = cfg["protected_target"]["position"]
    for r, want in zip(cfg["radars"], (1.5, 0.8, 2.5)):
        lin = lambda d: 10 ** (d / 10.0)
        d = ((r["position"][0] - tx) ** 2 + (r["position"][1] - ty) ** 2) ** 0.5
        unit = (lin(r["gt"]) * lin(r["gr"]) * r["wavelength"] ** 2 * cfg["protected_target"]["rcs"]
                / ((4 * np.pi) ** 3 * d ** 4 * lin(r["loss"]) * lin(r["latm"])))
        pn_w = 10 ** ((r["pn"] - 30) / 10)
        r["pt"] = float(f"{want * pn_w / (r['pulse_compression_gain'] * unit):.5g}")
    return cfg


SCENARIOS = {"default": scenario_default, "selftest": scenario_selftest, "active": scenario_active}


def f32_uniform(rng, size=None):
    """Uniform draws that are exactly representable in float32 (and < 1)."""
    u = rng.random(size).astype(np.float32)
    u = np.minimum(u, np.float32(1.0) - np.float32(2.0 ** -24))
    return u.astype(np.float64)


def gen_env():
    from simulation.environment import ElectromagneticEnvironment
    import simulation.environment as envmod  # noqa: F401

    out = {}
    for name, fn in SCENARIOS.items():
        cfg = fn()
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "sim.yaml")
            with open(path, "w") as f:
                yaml.safe_dump(cfg, f)
            rl = types.SimpleNamespace(episode_limit=7) if name == "active" else types.SimpleNamespace()
            with quiet():
                env = ElectromagneticEnvironment(rl, sim_config_path=path)
        R, J = env.num_radars, env.num_jammers
        A = 2 * R + 1
        rng = np.random.default_rng({"default": 11, "selftest": 12, "active": 13}[name])
        n_eps, T = (4, 60) if name != "active" else (6, 40)
        info0 = env.get_env_info()
        rec = {k: [] for k in ("act_d", "act_p", "noise", "reward", "r_d", "r_p", "r_j", "pd", "snr0",
                               "snr1", "tracking", "terminated", "step_count", "n_rng_draws")}
        states, obs_l, avail_l = [], [], []
        orig_rand = np.random.rand
        for ep in range(n_eps):
            with quiet():
                s0 = env.reset()
            states.append(s0)
            obs_l.append(np.stack(env.get_obs()))
            avail_l.append(np.stack(env.get_avail_actions()))
            for t in range(T):
                # actions: mostly valid, sometimes out of range / negative; power sometimes outside [0,1]
                act_d = rng.integers(0, A, size=J)
                weird = rng.random(J) < 0.08
                act_d = np.where(weird, rng.choice([-1, A, A + 2, 99], size=J), act_d)
                act_p = rng.random(J).astype(np.float32)
                edge = rng.random(J)
                act_p = np.where(edge < 0.05, np.float32(0.0), act_p)
                act_p = np.where((edge >= 0.05) & (edge < 0.10), np.float32(1.0), act_p)
                act_p = np.where((edge >= 0.10) & (edge < 0.13), np.float32(1.7), act_p)
                act_p = np.where((edge >= 0.13) & (edge < 0.16), np.float32(-0.4), act_p).astype(np.float32)
                u_radar = f32_uniform(rng, R)
                u_extra = f32_uniform(rng, J)
                fifo = list(u_radar) + list(u_extra)
                calls = [0]

                def fake_rand(*a, _fifo=fifo, _calls=calls):
                    assert not a
                    v = _fifo[_calls[0]]
                    _calls[0] += 1
                    return v

                np.random.rand = fake_rand
                try:
                    with quiet():
                        _, reward, term, info = env.step([(int(d), float(p)) for d, p in zip(act_d, act_p)])
                finally:
                    np.random.rand = orig_rand
                dec_jammers = [a["jammer_idx"] for a in info["jammer_actions"] if a["type"] == 0]
                assert calls[0] == R + len(dec_jammers)
                noise = np.concatenate([u_radar, f32_uniform(rng, J)])
                for k, j in enumerate(dec_jammers):
                    noise[R + j] = u_extra[k]
                rec["act_d"].append(act_d.astype(np.int32))
                rec["act_p"].append(act_p)
                rec["noise"].append(noise.astype(np.float32))
                rec["reward"].append(float(reward))
                rec["r_d"].append(float(info["r_d"]))
                rec["r_p"].append(float(info["r_p"]))
                rec["r_j"].append(float(info["r_j"]))
                rec["pd"].append(np.asarray(info["radar_pds"], dtype=np.float64))
                rec["snr0"].append(np.asarray(info["snr_no_jamming"], dtype=np.float64))
                rec["snr1"].append(np.asarray(info["snr_with_jamming"], dtype=np.float64))
                rec["tracking"].append(np.array([s["is_tracking"] for s in info["radar_states"]]))
                rec["terminated"].append(bool(term))
                rec["step_count"].append(env._step_count)
                rec["n_rng_draws"].append(calls[0])
        arrs = {k: np.asarray(v).reshape((n_eps, T) + np.asarray(v[0]).shape) for k, v in rec.items()}
        arrs["state0"] = np.stack(states)
        arrs["obs0"] = np.stack(obs_l)
        arrs["avail0"] = np.stack(avail_l)
        arrs["config_json"] = np.array(json.dumps(cfg))
        arrs["env_info_json"] = np.array(json.dumps(info0))
        arrs["episode_limit"] = np.array(env.episode_limit)
        # entity constants the reference derived (pins the dB conversions)
        arrs["radar_gt_lin"] = np.array([r.gt for r in env.radars])
        arrs["radar_pn_watts"] = np.array([r.pn_watts for r in env.radars])
        arrs["jammer_gj_lin"] = np.array([j.gj for j in env.jammers])
        tpos = env.protected_target_config["position"]
        arrs["echo_ps"] = np.array([r.calculate_echo_power(tpos, env.protected_target_config["rcs"]) for r in env.radars])
        arrs["pd_table_snr"] = np.array([0.0, 1e-5, 0.05, 0.2163, 0.5, 1.0, 2.0, 5.0, 80.0, 1e3])
        arrs["pd_table"] = np.array([env.radars[0].detection_probability(s) for s in arrs["pd_table_snr"]])
        np.savez_compressed(os.path.join(HERE, f"env_{name}.npz"), **arrs)
        print(f"env_{name}: {n_eps}x{T} steps, reward range [{arrs['reward'].min():.4f}, {arrs['reward'].max():.4f}], "
              f"pd range [{arrs['pd'].min():.4f}, {arrs['pd'].max():.4f}], r_j max {arrs['r_j'].max():.4f}")


GENERATORS = {"env": gen_env}

if __name__ == "__main__":
    which = sys.argv[1:] or list(GENERATORS)
    os.chdir(REF)  # the reference resolves config/ relative to the cwd
    for w in which:
        GENERATORS[w]()
