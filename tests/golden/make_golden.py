#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED reference.

Runs only in the build container, where the reference is mounted read-only at
/root/reference (it does not exist on the GPU box; the tests read the committed
.npz / .json files only).  Usage:

    python tests/golden/make_golden.py [env] [agent] [mixer] [learner] [replay] [api]
    MAKE_GOLDEN_NAMES=c3,c4 python tests/golden/make_golden.py agent mixer learner     (only these network configs)

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

REF = os.environ.get("MAKE_GOLDEN_REF", "/root/reference")     # (tests/test_oracle_live_reference.py points it at baseline/_ref)
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
    for name, fn in SCENARIOS.items():
        cfg = fn()
        rl = types.SimpleNamespace(episode_limit=7) if name == "active" else types.SimpleNamespace()
        n_eps, T = (4, 60) if name != "active" else (6, 40)
        arrs = record_env(cfg, rl, {"default": 11, "selftest": 12, "active": 13}[name], n_eps, T)
        np.savez_compressed(os.path.join(HERE, f"env_{name}.npz"), **arrs)
        print(f"env_{name}: {n_eps}x{T} steps, reward range [{arrs['reward'].min():.4f}, {arrs['reward'].max():.4f}], "
              f"pd range [{arrs['pd'].min():.4f}, {arrs['pd'].max():.4f}], r_j max {arrs['r_j'].max():.4f}")


def record_env(cfg, rl, seed, n_eps, T):
    """Drive ONE unmodified reference environment through n_eps episodes of T steps with seeded actions and injected
    uniform draws; returns everything the oracle / kernel checks compare against (the arrays of an env_*.npz)."""
    from simulation.environment import ElectromagneticEnvironment
    import simulation.environment as envmod  # noqa: F401

    if True:
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "sim.yaml")
            with open(path, "w") as f:
                yaml.safe_dump(cfg, f)
            with quiet():
                env = ElectromagneticEnvironment(rl, sim_config_path=path)
        R, J = env.num_radars, env.num_jammers
        A = 2 * R + 1
        rng = np.random.default_rng(seed)
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
        return arrs


def scenario_random(seed):
    """A random scenario for the live differential test: 1-5 radars, 1-4 jammers, random geometry, gains, losses,
    bandwidths, power ranges, radar types and threat levels; transmit powers scaled so that the un-jammed SNR lands
    between 0.3 and 4 (Pd between ~0.25 and ~1) and jamming powers that move it."""
    rng = np.random.default_rng(seed)
    R, J = int(rng.integers(1, 6)), int(rng.integers(1, 5))
    n_types = int(rng.integers(max(2, 1), 6))
    u = lambda lo, hi: float(np.round(rng.uniform(lo, hi), 4))
    tpos = [u(-50, 50), u(-50, 50)]
    radars = []
    for _ in range(R):
        ang, dist = rng.uniform(0, 2 * np.pi), rng.uniform(300, 3000)
        radars.append(dict(pt=1.0, gt=u(20, 35), gr=u(20, 35), wavelength=u(0.02, 0.1), rcs=u(0.5, 2.0), loss=u(2, 10), latm=u(0.5, 3),
                           pn=u(-95, -55), type_id=int(rng.integers(0, n_types)),
                           position=[float(np.round(tpos[0] + dist * np.cos(ang), 3)), float(np.round(tpos[1] + dist * np.sin(ang), 3))],
                           theta_m=u(1, 4), theta_a=u(0, 360), t_s=u(1, 8), pulse_compression_gain=u(20, 150),
                           anti_jamming_factor=u(1, 20), threat_level=u(0.4, 1.6)))
    jammers = []
    for _ in range(J):
        pmin = u(0, 2) if rng.random() < 0.5 else 0.0
        pmax = pmin if rng.random() < 0.15 else pmin + u(0.1, 20)
        jammers.append(dict(power=u(1, 10), gj=u(0, 25), loss=u(1, 6), latm=u(0.5, 3), bj=float(np.round(rng.uniform(0.5e6, 5e6), 0)),
                            position=[u(-200, 200), u(-200, 200)], power_min=pmin, power_max=float(np.round(pmax, 4))))
    cfg = {"radars": radars, "jammers": jammers, "protected_target": dict(position=tpos, rcs=u(0.5, 5.0)),
           "environment_params": dict(max_radar_types=n_types, rewards=dict(rd_min=u(-1.5, -1.0), rd_max=u(-0.9, -0.5),
                                                                             rp_min=u(-0.3, -0.1), rp_max=u(-0.05, -0.005)))}
    lin = lambda d: 10 ** (d / 10.0)
    for r in radars:
        d = ((r["position"][0] - tpos[0]) ** 2 + (r["position"][1] - tpos[1]) ** 2) ** 0.5
        unit = (lin(r["gt"]) * lin(r["gr"]) * r["wavelength"] ** 2 * cfg["protected_target"]["rcs"]
                / ((4 * np.pi) ** 3 * d ** 4 * lin(r["loss"]) * lin(r["latm"])))
        pn_w = 10 ** ((r["pn"] - 30) / 10)
        r["pt"] = float(f"{rng.uniform(0.3, 4.0) * pn_w / (r['pulse_compression_gain'] * unit):.5g}")
    return cfg


def gen_env_random():
    """Four random scenarios (seeds 300-303) committed as env_rand{0..3}.npz: they travel to the GPU box, where the live
    differential test (tests/test_oracle_live_reference.py) cannot run."""
    gen_live(HERE, 300, 4, prefix="env_rand")


def gen_live_nets(outdir, seed0, count):
    """`make_golden.py live_nets OUTDIR SEED COUNT`: agent / mixer / learner recordings of the reference at COUNT random
    sets of network dims (names live0 ..), into OUTDIR -- for tests/test_oracle_live_reference.py, not committed."""
    global HERE
    rng = np.random.default_rng(seed0)
    names = []
    for i in range(count):
        name = f"live{i}"
        S = int(rng.integers(5, 61))
        NET_CONFIGS[name] = dict(n_agents=int(rng.integers(1, 7)), n_actions=int(rng.integers(2, 17)), state_shape=S, obs_shape=S,
                                 rnn_hidden_dim=int(rng.choice([64, 128, 192])), actor_hidden_dim=int(rng.choice([64, 128])),
                                 mixing_embed_dim=int(rng.choice([16, 32, 64])), hyper_hidden_dim=int(rng.choice([32, 64, 128])))
        AGENT_SEEDS[name] = (int(rng.integers(1, 1 << 30)), int(rng.integers(1, 1 << 30)), int(rng.integers(2, 8)))
        MIXER_SEEDS[name] = int(rng.integers(1, 1 << 30))
        LEARNER_CONFIGS[name] = dict(net=name, B=int(rng.integers(2, 6)), T=int(rng.integers(3, 8)), ragged=bool(rng.random() < 0.5),
                                     lr=float(rng.choice([5e-6, 1e-4, 1e-3])), interval=int(rng.integers(1, 4)), steps=3)
        names.append(name)
    os.environ["MAKE_GOLDEN_NAMES"] = ",".join(names)
    HERE = outdir
    gen_agent()
    gen_mixer()
    gen_learner()


def gen_live_replay(outdir, seed0, count):
    """`make_golden.py live_replay OUTDIR SEED COUNT`: random rings (capacity, dims, episode lengths, sampling points) driven
    through the reference EpisodeReplayBuffer, the sampled indices drawn by the reference's own np.random.choice."""
    for i in range(count):
        rng = np.random.default_rng(seed0 + i)
        cap, T = int(rng.integers(2, 12)), int(rng.integers(2, 10))
        n_eps = int(rng.integers(cap, 3 * cap + 2))
        lens = [int(T if rng.random() < 0.5 else rng.integers(1, T + 1)) for _ in range(n_eps)]
        sample_at = sorted(set(int(x) for x in rng.integers(0, n_eps, size=4)) | {n_eps - 1})
        gen_replay(os.path.join(outdir, f"replay_live{i}.npz"), seed=seed0 + i, cap=cap, T=T, Nn=int(rng.integers(1, 5)),
                   A=int(rng.integers(2, 9)), S=int(rng.integers(3, 20)), H=int(rng.choice([8, 16, 64])), lens=lens, sample_at=sample_at,
                   n_sample=int(rng.integers(1, cap + 2)), seeded_draws=True)


def gen_live(outdir, seed0, count, prefix="env_live_"):
    """`make_golden.py live OUTDIR SEED COUNT`: COUNT random scenarios recorded from the reference into OUTDIR (not
    committed: tests/test_oracle_live_reference.py calls this in a subprocess when a copy of the reference is present)."""
    for i in range(count):
        cfg = scenario_random(seed0 + i)
        limit = int(np.random.default_rng(seed0 + i).integers(5, 30))
        arrs = record_env(cfg, types.SimpleNamespace(episode_limit=limit), 1000 + seed0 + i, 3, limit + 3)
        np.savez_compressed(os.path.join(outdir, f"{prefix}{i}.npz"), **arrs)
        print(f"{prefix}{i}: R={len(cfg['radars'])} J={len(cfg['jammers'])} limit={limit} pd range [{arrs['pd'].min():.3f}, {arrs['pd'].max():.3f}] "
              f"r_j max {arrs['r_j'].max():.3f} terminated {int(arrs['terminated'].sum())}")


# ------------------------------------------------------------------ networks
def rl_args(**kw):
    base = dict(n_agents=2, n_actions=5, state_shape=24, obs_shape=24, rnn_hidden_dim=128, actor_hidden_dim=128,
                mixing_embed_dim=64, hyper_hidden_dim=128, epsilon_start=1.0, epsilon_finish=0.05,
                epsilon_anneal_time=100000, gamma=0.99, lr=5e-6, grad_norm_clip=1.0, target_update_interval=200,
                use_cuda=False, device="cpu", batch_size=32, buffer_size=16, episode_limit=100)
    base.update(kw)
    return types.SimpleNamespace(**base)


NET_CONFIGS = {
    "c1": dict(),
    "small": dict(n_agents=3, n_actions=7, state_shape=39, obs_shape=39, rnn_hidden_dim=64, actor_hidden_dim=64,
                  mixing_embed_dim=32, hyper_hidden_dim=64),
    # BASELINE.json configs[2]: 8 jammers x 16 radars x 4 target types -> 33 actions, state = 16 x (6 + 4) + 2 x 8 = 176
    "c3": dict(n_agents=8, n_actions=33, state_shape=176, obs_shape=176),
    # BASELINE.json configs[3]: the learner stress dims (GRU 256, mixer embed 128)
    "c4": dict(rnn_hidden_dim=256, mixing_embed_dim=128),
}
AGENT_SEEDS = {"c1": (42, 100, 6), "small": (7, 101, 5), "c3": (11, 102, 4), "c4": (12, 103, 6)}      # torch seed, numpy seed, B
MIXER_SEEDS = {"c1": 3, "small": 4, "c3": 5, "c4": 6}


def wanted(name):
    only = os.environ.get("MAKE_GOLDEN_NAMES")
    return only is None or name in only.split(",")


def realistic_obs(rng, shape):
    """Half unit normal, half with the magnitudes the real state vector has (pt ~ 300, pos ~ 400)."""
    x = rng.standard_normal(shape).astype(np.float32)
    scale = np.where(rng.random(shape[-1]) < 0.5, 1.0, rng.choice([5.0, 50.0, 300.0], size=shape[-1])).astype(np.float32)
    return x * scale


def gen_agent():
    import torch
    from core.mac import BasicMAC
    for name, kw in NET_CONFIGS.items():
        if not wanted(name):
            continue
        args = rl_args(**kw)
        torch.manual_seed(AGENT_SEEDS[name][0])
        with quiet():
            mac = BasicMAC(args.obs_shape, args)
        sd = {k: v.detach().clone() for k, v in mac.agent.state_dict().items()}
        rng = np.random.default_rng(AGENT_SEEDS[name][1])
        B, Nn, O, A, H = AGENT_SEEDS[name][2], args.n_agents, args.obs_shape, args.n_actions, args.rnn_hidden_dim
        steps = 4
        out = {f"sd.{k}": v.numpy() for k, v in sd.items()}
        out["n_params"] = np.array(sum(v.numel() for v in sd.values()))
        h0 = (rng.standard_normal((B * Nn, H)) * 0.5).astype(np.float32)
        rec = {k: [] for k in ("obs", "avail", "u", "rand_actions", "t_env", "eps", "actions", "actions_test",
                               "power", "power_test", "hidden", "q", "params")}
        orig_rand_like, orig_multinomial = torch.rand_like, torch.multinomial
        mac.hidden_states = torch.from_numpy(h0.copy())
        for t in range(steps):
            obs = realistic_obs(rng, (B, Nn, O))
            avail = (rng.random((B, Nn, A)) < 0.7).astype(np.int64)
            avail[..., 0] = np.where(avail.sum(-1) == 0, 1, avail[..., 0])
            if t == 0:
                avail[:] = 1
            u = rng.random((B, Nn)).astype(np.float32)
            t_env = int(rng.integers(0, 120000))
            # a random *available* action per agent (what torch.multinomial would draw from)
            ra = np.array([[rng.choice(np.flatnonzero(avail[b, n])) for n in range(Nn)] for b in range(B)], dtype=np.int64)
            h_before = mac.hidden_states.clone()
            torch.rand_like = lambda x, _u=u: torch.from_numpy(_u.copy())
            torch.multinomial = lambda w, num_samples, _ra=ra: torch.from_numpy(_ra.reshape(-1, 1).copy())
            try:
                a_test, p_test = mac.select_actions(torch.from_numpy(obs), torch.from_numpy(avail), t_env, test_mode=True)
                mac.hidden_states = h_before.clone()
                a, p = mac.select_actions(torch.from_numpy(obs), torch.from_numpy(avail), t_env, test_mode=False)
            finally:
                torch.rand_like, torch.multinomial = orig_rand_like, orig_multinomial
            # all-action Q and params through the reference agent directly
            x = torch.from_numpy(obs).reshape(B * Nn, O)
            h2 = mac.agent.forward(x, h_before)
            params = mac.agent.actor_forward(x)
            q = torch.stack([mac.agent.get_q_value_for_action(h2, torch.full((B * Nn, 1), ai, dtype=torch.long),
                                                              params[:, ai:ai + 1]).squeeze(1) for ai in range(A)], 1)
            assert torch.equal(h2.detach(), mac.hidden_states)
            rec["obs"].append(obs); rec["avail"].append(avail); rec["u"].append(u); rec["rand_actions"].append(ra)
            rec["t_env"].append(t_env); rec["eps"].append(mac.action_selector.epsilon)
            rec["actions"].append(a.numpy()); rec["actions_test"].append(a_test.numpy())
            rec["power"].append(p.detach().numpy()); rec["power_test"].append(p_test.detach().numpy())
            rec["hidden"].append(mac.hidden_states.numpy().copy()); rec["q"].append(q.detach().numpy())
            rec["params"].append(params.detach().numpy())
        out.update({k: np.stack(v) for k, v in rec.items()})
        out["h0"] = h0
        out["args_json"] = np.array(json.dumps(vars(args)))
        np.savez_compressed(os.path.join(HERE, f"agent_{name}.npz"), **out)
        print(f"agent_{name}: params {int(out['n_params'])}, q range [{out['q'].min():.3f}, {out['q'].max():.3f}]")


def gen_mixer():
    import torch
    from core.networks import QMixer
    for name, kw in NET_CONFIGS.items():
        if not wanted(name):
            continue
        args = rl_args(**kw)
        torch.manual_seed(MIXER_SEEDS[name])
        mixer = QMixer(args)
        rng = np.random.default_rng(200)
        Rr = 37
        q = (rng.standard_normal((Rr, args.n_agents)) * 2).astype(np.float32)
        s = realistic_obs(rng, (Rr, args.state_shape))
        out = {f"sd.{k}": v.detach().numpy() for k, v in mixer.state_dict().items()}
        out["n_params"] = np.array(sum(v.numel() for v in mixer.state_dict().values()))
        qt = torch.from_numpy(q).requires_grad_(True)
        y = mixer(qt.view(1, Rr, args.n_agents), torch.from_numpy(s).view(1, Rr, -1)).view(Rr)
        # gradient of sum(y * w) w.r.t. everything: pins the backward pass
        w = rng.standard_normal(Rr).astype(np.float32)
        (y * torch.from_numpy(w)).sum().backward()
        out.update(q=q, s=s, y=y.detach().numpy(), w=w, dq=qt.grad.numpy())
        out.update({f"grad.{k}": v.grad.numpy() for k, v in mixer.named_parameters()})
        out["args_json"] = np.array(json.dumps(vars(args)))
        np.savez_compressed(os.path.join(HERE, f"mixer_{name}.npz"), **out)
        print(f"mixer_{name}: params {int(out['n_params'])}, y range [{out['y'].min():.3f}, {out['y'].max():.3f}]")


def synthetic_batch(rng, args, B, T, ragged):
    """A sampled-batch dict in the reference's layout (utils/replay_buffer.py:153-214)."""
    Nn, A, S, O, H = args.n_agents, args.n_actions, args.state_shape, args.obs_shape, args.rnn_hidden_dim
    lens = np.full(B, T)
    if ragged:
        lens[1:] = rng.integers(2, T + 1, size=B - 1)
        lens[0] = T
    filled = (np.arange(T)[None, :] < lens[:, None])[..., None]
    terminated = ~filled
    for b in range(B):
        if ragged and lens[b] < T:
            terminated[b, lens[b] - 1, 0] = True
    batch = {
        "state": realistic_obs(rng, (B, T + 1, S)),
        "obs": realistic_obs(rng, (B, T + 1, Nn, O)),
        "actions_discrete": rng.integers(0, A, size=(B, T, Nn, 1)).astype(np.int32),
        "actions_continuous": rng.random((B, T, Nn, 1)).astype(np.float32),
        "avail_actions": np.ones((B, T + 1, Nn, A), dtype=np.int64),
        "reward": rng.standard_normal((B, T, 1)).astype(np.float32),
        "terminated": terminated,
        "filled": filled,
        "hidden_state": (rng.standard_normal((B, T + 1, Nn, H)) * 0.5).astype(np.float32),
        "max_seq_len": int(lens.max()),
    }
    return batch


LEARNER_CONFIGS = {
    "c1": dict(net="c1", B=4, T=7, ragged=False, lr=5e-6, interval=2, steps=3),
    "small_fastlr": dict(net="small", B=5, T=6, ragged=True, lr=1e-3, interval=2, steps=4),
    "c3": dict(net="c3", B=3, T=5, ragged=True, lr=5e-6, interval=2, steps=2),
    "c4": dict(net="c4", B=4, T=6, ragged=False, lr=1e-4, interval=1, steps=2),
}


def gen_learner():
    import torch
    from core.mac import BasicMAC
    from core.qmix import QMixLearner
    for name, c in LEARNER_CONFIGS.items():
        if not wanted(name):
            continue
        args = rl_args(**NET_CONFIGS[c["net"]], lr=c["lr"], target_update_interval=c["interval"])
        torch.manual_seed(42)
        with quiet():
            mac = BasicMAC(args.obs_shape, args)
            learner = QMixLearner(mac, args)
        rng = np.random.default_rng(300)
        out = {f"agent0.{k}": v.detach().numpy().copy() for k, v in mac.agent.state_dict().items()}
        out.update({f"mixer0.{k}": v.detach().numpy().copy() for k, v in learner.eval_qmix_net.state_dict().items()})
        for step in range(c["steps"]):
            batch = synthetic_batch(rng, args, c["B"], c["T"], c["ragged"])
            stats = learner.train(batch, {})
            for k, v in batch.items():
                out[f"step{step}.batch.{k}"] = np.asarray(v)
            out[f"step{step}.stats"] = np.array([stats["loss"], stats["grad_norm"], stats["eval_qtot_avg"], stats["target_qtot_avg"]])
            for k, p_ in mac.agent.named_parameters():
                out[f"step{step}.agent_has_grad.{k}"] = np.array(p_.grad is not None)
                if p_.grad is not None:
                    out[f"step{step}.agent_grad.{k}"] = p_.grad.numpy().copy()      # after clipping
            for k, p_ in learner.eval_qmix_net.named_parameters():
                out[f"step{step}.mixer_grad.{k}"] = p_.grad.numpy().copy()
            # only the Q-head ever changes: store it, assert the rest is frozen
            a0 = {k[len("agent0."):]: v for k, v in out.items() if k.startswith("agent0.")}
            frozen = all(np.array_equal(a0[k], v.detach().numpy()) for k, v in mac.agent.state_dict().items()
                         if not k.startswith("fc2_q_head"))
            out[f"step{step}.agent_untrained_unchanged"] = np.array(frozen)
            assert frozen
            out.update({f"step{step}.agent.{k}": v.detach().numpy().copy() for k, v in mac.agent.state_dict().items()
                        if k.startswith("fc2_q_head")})
            out.update({f"step{step}.mixer.{k}": v.detach().numpy().copy() for k, v in learner.eval_qmix_net.state_dict().items()})
            out.update({f"step{step}.tgt_agent.{k}": v.detach().numpy().copy()
                        for k, v in learner.target_mac.agent.state_dict().items() if k.startswith("fc2_q_head")})
            out.update({f"step{step}.tgt_mixer.{k}": v.detach().numpy().copy() for k, v in learner.target_qmix_net.state_dict().items()})
            print(f"learner_{name} step {step}: {stats}")
        out["args_json"] = np.array(json.dumps(vars(args)))
        out["n_steps"] = np.array(c["steps"])
        np.savez_compressed(os.path.join(HERE, f"learner_{name}.npz"), **out)


def gen_replay(path=None, seed=400, cap=5, T=6, Nn=2, A=5, S=8, H=16, lens=(6, 6, 3, 6, 5, 6, 6), sample_at=(2, 6), n_sample=3,
               seeded_draws=False):
    """Defaults: the committed replay.npz (7 episodes into a ring of 5 -> wraps; the sampled indices injected).
    seeded_draws=True (live differential test): sample() draws its indices itself after np.random.seed(recorded seed)."""
    from utils.replay_buffer import EpisodeReplayBuffer
    args = rl_args(buffer_size=cap, episode_limit=T, n_agents=Nn, n_actions=A, state_shape=S, obs_shape=S, rnn_hidden_dim=H)
    with quiet():
        buf = EpisodeReplayBuffer(args)
    rng = np.random.default_rng(seed)
    out = {"args_json": np.array(json.dumps(vars(args)))}
    for i, L in enumerate(lens):
        ep = {
            "state": [rng.standard_normal((L + 1, S)).astype(np.float32)],
            "obs": [rng.standard_normal((L + 1, Nn, S)).astype(np.float32)],
            "actions_discrete": [rng.integers(0, A, size=(L, Nn, 1)).astype(np.int32)],
            "actions_continuous": [rng.random((L, Nn, 1)).astype(np.float32)],
            "avail_actions": [np.ones((L + 1, Nn, A), dtype=np.int64)],
            "reward": [rng.standard_normal((L, 1)).astype(np.float32)],
            "terminated": [np.arange(L)[:, None] == L - 1],
            "hidden_state": [rng.standard_normal((L + 1, Nn, H)).astype(np.float32)],
        }
        for k, v in ep.items():
            out[f"ep{i}.{k}"] = v[0]
        buf.store_episode(ep)
        out[f"after{i}.current_index"] = np.array(buf.current_index)
        out[f"after{i}.current_size"] = np.array(buf.current_size)
        if i in sample_at and seeded_draws:
            np_seed = int(rng.integers(0, 1 << 31))
            n_draw = min(n_sample, buf.current_size)
            np.random.seed(np_seed)
            idx = np.random.choice(buf.current_size, n_draw, replace=False)      # what sample() is about to draw (replay_buffer.py:178)
            np.random.seed(np_seed)
            with quiet():
                b = buf.sample(n_draw)
            out[f"sample{i}.np_seed"] = np.array(np_seed)
            out[f"sample{i}.indices"] = idx
            for k, v in b.items():
                out[f"sample{i}.{k}"] = np.asarray(v)
        elif i in sample_at:
            idx = rng.permutation(buf.current_size)[:n_sample]
            orig = np.random.choice
            np.random.choice = lambda n, k, replace=False, _idx=idx: _idx
            try:
                with quiet():
                    b = buf.sample(n_sample)
            finally:
                np.random.choice = orig
            out[f"sample{i}.indices"] = idx
            for k, v in b.items():
                out[f"sample{i}.{k}"] = np.asarray(v)
    out["n_eps"] = np.array(len(lens))
    np.savez_compressed(path or os.path.join(HERE, "replay.npz"), **out)
    print("replay: stored", len(lens), "episodes; final index/size", buf.current_index, buf.current_size)


def gen_api():
    """Public method signatures of the reference classes on the hot-path boundary."""
    from simulation.environment import ElectromagneticEnvironment
    from core.mac import BasicMAC
    from core.networks import RNNAgent, QMixer
    from core.qmix import QMixLearner
    from utils.replay_buffer import EpisodeReplayBuffer
    from utils.action_selectors import EpsilonGreedyActionSelector
    from runners.episode_runner import EpisodeRunner, EpisodeBatch
    api = {}
    for cls in (ElectromagneticEnvironment, BasicMAC, RNNAgent, QMixer, QMixLearner, EpisodeReplayBuffer,
                EpsilonGreedyActionSelector, EpisodeRunner, EpisodeBatch):
        methods = {}
        for mname, fn in inspect.getmembers(cls, predicate=inspect.isfunction):
            if mname.startswith("_") and mname != "__init__":
                continue
            if fn.__qualname__.split(".")[0] != cls.__name__:
                continue   # inherited from nn.Module
            methods[mname] = [p_.name for p_ in inspect.signature(fn).parameters.values()]
        api[cls.__name__] = methods
    with open(os.path.join(HERE, "reference_api.json"), "w") as f:
        json.dump(api, f, indent=1, sort_keys=True)
    print("api:", {k: len(v) for k, v in api.items()})


GENERATORS = {"env": gen_env, "env_random": gen_env_random, "agent": gen_agent, "mixer": gen_mixer, "learner": gen_learner,
              "replay": gen_replay, "api": gen_api}

if __name__ == "__main__":
    which = sys.argv[1:] or list(GENERATORS)
    os.chdir(REF)  # the reference resolves config/ relative to the cwd
    if which[0] in ("live", "live_nets", "live_replay"):
        {"live": gen_live, "live_nets": gen_live_nets, "live_replay": gen_live_replay}[which[0]](which[1], int(which[2]), int(which[3]))
        sys.exit(0)
    for w in which:
        GENERATORS[w]()
