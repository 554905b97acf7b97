"""Scenario tables for the batched electromagnetic environment.

A *raw spec* is the reference's scenario YAML (config/simulation_config.yaml)
with a leading ``n_envs`` axis on every entity parameter, kept in the YAML's own
units (gains / losses / noise in dB).  ``ScenarioTables`` turns it into the
float64 struct-of-arrays tables the step kernel reads from HBM.

Reference behaviour mirrored here:
  * required keys and error types of simulation/environment.py:45-73,123-206
    (ValueError for missing top-level keys / bad type_id, KeyError for missing
    entity parameters)
  * counts come from the entity lists, ``episode_limit`` from the RL config with
    the hard-coded fallback 100 (environment.py:82-84)
  * jammer ``power`` in the YAML is ignored; power_min/max default to 0/100 W
    (environment.py:184-190)
  * dB -> linear conversions of core/radar.py:10-19 and core/jammer.py:44-47
"""
from __future__ import annotations

import numpy as np
import yaml

RADAR_KEYS = ("pt", "gt", "gr", "wavelength", "rcs", "loss", "latm", "pn", "type_id",
              "position", "theta_m", "theta_a", "t_s", "pulse_compression_gain",
              "anti_jamming_factor")
JAMMER_KEYS = ("gj", "loss", "latm", "bj", "position")


def load_sim_config(path):
    """Parse and validate a scenario YAML the way the reference constructor does."""
    with open(path, "r") as f:
        cfg = yaml.safe_load(f)
    if not cfg or "radars" not in cfg or "jammers" not in cfg:
        raise ValueError(f"Simulation config file {path} is missing required 'radars' or 'jammers' keys.")
    if "protected_target" not in cfg and "protected_targets" not in cfg:
        raise ValueError(f"Simulation config file {path} is missing required 'protected_target' key.")
    return cfg


def spec_from_config(cfg, n_envs=1, num_radars=None, num_jammers=None, episode_limit=100):
    """Raw spec (every env identical) from a parsed scenario YAML dict."""
    radars = list(cfg.get("radars", []))
    jammers = list(cfg.get("jammers", []))
    if num_radars is not None:
        radars = radars[:num_radars]
    if num_jammers is not None:
        jammers = jammers[:num_jammers]
    envp = cfg.get("environment_params", {}) or {}
    types = int(envp.get("max_radar_types", 4))
    rew = envp.get("rewards", {}) or {}
    if "protected_targets" in cfg:
        targets = list(cfg["protected_targets"])
    else:
        targets = [cfg["protected_target"]]
    for t in targets:
        if "position" not in t or "rcs" not in t:
            raise ValueError("Protected target config must contain 'position' and 'rcs'.")
    for i, r in enumerate(radars):
        for k in RADAR_KEYS:
            if k not in r:
                raise KeyError(f"Radar config {i} missing required parameter: {k}")
        if not (0 <= r.get("type_id", -1) < types):
            raise ValueError(f"Radar config {i} invalid type_id")
    for i, j in enumerate(jammers):
        for k in JAMMER_KEYS:
            if k not in j:
                raise KeyError(f"Jammer config {i} missing required parameter: {k}")

    def rep(vals, dtype=np.float64):
        a = np.asarray(vals, dtype=dtype)
        return np.broadcast_to(a[None], (n_envs,) + a.shape).copy()

    rad = {k: rep([r[k] for r in radars]) for k in RADAR_KEYS if k != "type_id"}
    rad["type_id"] = rep([r["type_id"] for r in radars], np.int64)
    rad["threat_level"] = rep([r.get("threat_level", 1.0) for r in radars])
    jam = {k: rep([j[k] for j in jammers]) for k in JAMMER_KEYS}
    jam["power_min"] = rep([j.get("power_min", 0.0) for j in jammers])
    jam["power_max"] = rep([j.get("power_max", 100.0) for j in jammers])
    tgt = {"position": rep([t["position"] for t in targets]),
           "rcs": rep([t["rcs"] for t in targets])}
    env = {"max_radar_types": types,
           "rd_min": rew.get("rd_min", -1.2), "rd_max": rew.get("rd_max", -0.8),
           "rp_min": rew.get("rp_min", -0.1), "rp_max": rew.get("rp_max", -0.01),
           "episode_limit": int(episode_limit)}
    return {"radars": rad, "jammers": jam, "targets": tgt, "env": env}


def default_config_dict():
    """The shipped default scenario (config/simulation_config.yaml of the reference:
    2 radars, 2 jammers, one protected target at the origin), as a parsed dict so
    benchmarks and GPU tests do not need the reference tree."""
    return {
        "radars": [
            dict(type_id=1, pt=300.0, gt=30, gr=30, wavelength=0.03, rcs=1.0, loss=10, latm=2, pn=3,
                 position=[400.0, 0.0], theta_m=2.0, theta_a=0.0, t_s=5.0,
                 pulse_compression_gain=100.0, anti_jamming_factor=10.0, threat_level=0.8),
            dict(type_id=2, pt=180.0, gt=25, gr=25, wavelength=0.03, rcs=1.0, loss=10, latm=2, pn=3,
                 position=[-400.0, 0.0], theta_m=1.8, theta_a=180.0, t_s=4.0,
                 pulse_compression_gain=120.0, anti_jamming_factor=15.0, threat_level=1.2),
        ],
        "jammers": [
            dict(power=1000, gj=20, loss=5, latm=2, bj=10, position=[50.0, 50.0]),
            dict(power=1000, gj=20, loss=5, latm=2, bj=10, position=[-50.0, -50.0]),
        ],
        "protected_target": dict(position=[0, 0], rcs=1.0),
        "environment_params": dict(max_radar_types=4,
                                   rewards=dict(rd_min=-1.2, rd_max=-0.8, rp_min=-0.1, rp_max=-0.01)),
    }


def default_spec(n_envs=1, episode_limit=100):
    """S-default of SURVEY.md section 8d: the shipped scenario replicated."""
    return spec_from_config(default_config_dict(), n_envs=n_envs, episode_limit=episode_limit)


def hetero_spec(n_envs, seed=1234, active=False, episode_limit=100):
    """S-hetero / S-active: default parameter ranges, entity positions ~ U[-500,500]^2
    per env.  With ``active`` the transmit power is rescaled per radar so that
    Ga*Ps/Pn ~ U[0.1, 3] and Pd spans 0.13 .. 0.99 instead of the frozen 0.103."""
    spec = default_spec(n_envs, episode_limit)
    rng = np.random.default_rng(seed)
    R = spec["radars"]["pt"].shape[1]
    J = spec["jammers"]["gj"].shape[1]
    spec["radars"]["position"] = rng.uniform(-500, 500, size=(n_envs, R, 2))
    spec["jammers"]["position"] = rng.uniform(-500, 500, size=(n_envs, J, 2))
    spec["targets"]["position"] = rng.uniform(-20, 20, size=(n_envs, 1, 2))
    spec["radars"]["threat_level"] = rng.uniform(0.7, 1.3, size=(n_envs, R))
    spec["radars"]["type_id"] = rng.integers(0, 4, size=(n_envs, R)).astype(np.int64)
    if active:
        rad = spec["radars"]
        lin = lambda d: 10.0 ** (np.asarray(d, dtype=np.float64) / 10.0)
        d = np.linalg.norm(rad["position"][:, :, None, :] - spec["targets"]["position"][:, None, :, :], axis=-1)[..., 0]
        unit = (lin(rad["gt"]) * lin(rad["gr"]) * rad["wavelength"] ** 2 * spec["targets"]["rcs"][:, :1]
                / ((4 * np.pi) ** 3 * np.maximum(d, 1e-6) ** 4 * lin(rad["loss"]) * lin(rad["latm"])))
        pn = 10.0 ** ((rad["pn"] - 30.0) / 10.0)
        want = rng.uniform(0.1, 3.0, size=(n_envs, R))
        rad["pt"] = want * pn / (rad["pulse_compression_gain"] * unit)
        # keep the jammers relevant: received jamming power comparable with the noise floor
        spec["jammers"]["power_max"] = np.full((n_envs, J), 1e-2)
    return spec


def scaled_spec(n_envs, n_jammers=8, n_radars=16, n_targets=4, seed=1234, episode_limit=100, active=True):
    """S-scaled (BASELINE config 3): radars cycle the two default templates with
    type_id in {0..3} on a 400 m ring, jammers on a 70 m ring, targets ~ U[-20,20]^2."""
    base = default_config_dict()
    rng = np.random.default_rng(seed)
    radars = []
    for r in range(n_radars):
        t = dict(base["radars"][r % 2])
        ang = 2 * np.pi * r / n_radars
        t["position"] = [400.0 * np.cos(ang), 400.0 * np.sin(ang)]
        t["type_id"] = r % 4
        t["theta_a"] = float(np.degrees(ang))
        t["threat_level"] = 0.8 + 0.4 * (r % 5) / 4.0
        radars.append(t)
    jammers = []
    for j in range(n_jammers):
        t = dict(base["jammers"][j % 2])
        ang = 2 * np.pi * j / n_jammers + 0.1
        t["position"] = [70.0 * np.cos(ang), 70.0 * np.sin(ang)]
        jammers.append(t)
    cfg = dict(base, radars=radars, jammers=jammers)
    cfg.pop("protected_target")
    cfg["protected_targets"] = [dict(position=[0.0, 0.0], rcs=1.0) for _ in range(n_targets)]
    spec = spec_from_config(cfg, n_envs=n_envs, episode_limit=episode_limit)
    spec["targets"]["position"] = rng.uniform(-20, 20, size=(n_envs, n_targets, 2))
    spec["targets"]["rcs"] = rng.uniform(0.5, 2.0, size=(n_envs, n_targets))
    if active:
        rad = spec["radars"]
        lin = lambda d: 10.0 ** (np.asarray(d, dtype=np.float64) / 10.0)
        unit = (lin(rad["gt"]) * lin(rad["gr"]) * rad["wavelength"] ** 2
                / ((4 * np.pi) ** 3 * 400.0 ** 4 * lin(rad["loss"]) * lin(rad["latm"])))
        pn = 10.0 ** ((rad["pn"] - 30.0) / 10.0)
        want = rng.uniform(0.1, 3.0, size=(n_envs, n_radars))
        rad["pt"] = want * pn / (rad["pulse_compression_gain"] * unit)
        spec["jammers"]["power_max"] = np.full((n_envs, n_jammers), 1e-2)
    return spec


class ScenarioTables:
    """Float64 struct-of-arrays tables, ``[param][env]`` with the env index fastest.

    Layout (one float64 array ``data`` of ``n_rows * n_envs`` elements; row ``p``
    of entity ``e`` lives at ``data[(base + e * stride + p) * n_envs + env]``):

      radar r  (RADAR_ROWS = 16 rows): 0 pt, 1 gt_lin, 2 gr_lin, 3 wavelength,
               4 loss_lin, 5 latm_lin, 6 pn_watts, 7 Ga, 8 D, 9 threat,
               10 pos_x, 11 pos_y, 12 theta_m, 13 theta_a, 14 t_s, 15 type_id
      jammer j (JAMMER_ROWS = 8 rows): 0 gj_lin, 1 loss_lin, 2 latm_lin, 3 bj,
               4 pos_x, 5 pos_y, 6 power_min, 7 power_max
      target k (TARGET_ROWS = 3 rows): 0 pos_x, 1 pos_y, 2 rcs
    """
    RADAR_ROWS, JAMMER_ROWS, TARGET_ROWS = 16, 8, 3

    def __init__(self, spec):
        rad, jam, tgt, env = spec["radars"], spec["jammers"], spec["targets"], spec["env"]
        f = lambda a: np.asarray(a, dtype=np.float64)
        self.n_envs, self.R = f(rad["pt"]).shape
        self.J = f(jam["gj"]).shape[1]
        self.K = f(tgt["rcs"]).shape[1]
        self.types = int(env["max_radar_types"])
        tid = np.asarray(rad["type_id"])
        if np.any(tid < 0) or np.any(tid >= self.types):
            raise ValueError("Radar config invalid type_id")
        self.rd_min, self.rd_max = float(env["rd_min"]), float(env["rd_max"])
        self.rp_min, self.rp_max = float(env["rp_min"]), float(env["rp_max"])
        self.episode_limit = int(env["episode_limit"])
        self.state_dim = self.R * (6 + self.types) + 2 * self.J
        self.n_actions = 2 * self.R + 1
        lin = _db_to_linear
        rows = []
        for r in range(self.R):
            rows += [f(rad["pt"])[:, r], lin(rad["gt"])[:, r], lin(rad["gr"])[:, r],
                     f(rad["wavelength"])[:, r], lin(rad["loss"])[:, r], lin(rad["latm"])[:, r],
                     _noise_watts(rad["pn"])[:, r], f(rad["pulse_compression_gain"])[:, r],
                     f(rad["anti_jamming_factor"])[:, r], f(rad["threat_level"])[:, r],
                     f(rad["position"])[:, r, 0], f(rad["position"])[:, r, 1],
                     f(rad["theta_m"])[:, r], f(rad["theta_a"])[:, r], f(rad["t_s"])[:, r],
                     tid[:, r].astype(np.float64)]
        for j in range(self.J):
            rows += [lin(jam["gj"])[:, j], lin(jam["loss"])[:, j], lin(jam["latm"])[:, j],
                     f(jam["bj"])[:, j], f(jam["position"])[:, j, 0], f(jam["position"])[:, j, 1],
                     f(jam["power_min"])[:, j], f(jam["power_max"])[:, j]]
        for k in range(self.K):
            rows += [f(tgt["position"])[:, k, 0], f(tgt["position"])[:, k, 1], f(tgt["rcs"])[:, k]]
        self.data = np.ascontiguousarray(np.stack(rows, axis=0))      # [n_rows, n_envs]
        self.n_rows = self.data.shape[0]

    @property
    def radar_base(self):
        return 0

    @property
    def jammer_base(self):
        return self.R * self.RADAR_ROWS

    @property
    def target_base(self):
        return self.R * self.RADAR_ROWS + self.J * self.JAMMER_ROWS


def _db_to_linear(x):
    """10 ** (dB / 10) with python-float pow per element, as utils/math_utils.py:3-8
    evaluates it (keeps the tables bit-identical to the reference's entities)."""
    a = np.asarray(x, dtype=np.float64)
    uniq, inv = np.unique(a.ravel(), return_inverse=True)
    out = np.array([10 ** (float(v) / 10.0) for v in uniq], dtype=np.float64)
    return out[inv].reshape(a.shape)


def _noise_watts(pn_db):
    """dBm -> W as core/radar.py:19 does."""
    a = np.asarray(pn_db, dtype=np.float64)
    uniq, inv = np.unique(a.ravel(), return_inverse=True)
    out = np.array([10 ** ((float(v) - 30) / 10) for v in uniq], dtype=np.float64)
    return out[inv].reshape(a.shape)
