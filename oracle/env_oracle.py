"""NumPy float64 oracle of the electromagnetic environment step.  TEST INFRASTRUCTURE.

Vectorised over ``n_envs`` independent episodes; at ``n_envs == 1`` and one
protected target it is the reference's single-instance environment.

Follows (reference paths relative to /root/reference):
  * simulation/environment.py:221-477  step(): action decode, jammer power
    bookkeeping, Friis jamming power, echo power, SNR with/without jamming,
    Albersheim Pd, Monte-Carlo detection, r_d / r_p / r_j, termination
  * simulation/environment.py:479-551  get_state / get_obs / get_avail_actions
  * core/radar.py:10-33 (dB -> linear, pn_watts), 35-60 (echo power),
    67-82 (detection probability), 90-117 (memoryless track state)
  * core/jammer.py:24-54 (dB -> linear), 56-98 (received power)
  * utils/math_utils.py:3-8, 40-42 ; utils/state_utils.py:3-28

Parity status: PINNED against the unmodified reference run in the build
container (tests/golden/env_*.npz, produced by tests/golden/make_golden.py).

Extensions that do not exist in the reference and are
therefore pinned only by this restatement: K > 1 protected targets, the
jamming-to-signal ratio in dB and the networked detection probability.  With
K == 1 every reference-defined output is unchanged by them.

Scenario input format ("raw spec", shared *format* with the product package,
independent code): a dict
    radars : {pt, gt, gr, wavelength, rcs, loss, latm, pn, type_id, position,
              theta_m, theta_a, t_s, pulse_compression_gain,
              anti_jamming_factor, threat_level}     arrays [n, R] ([n, R, 2])
    jammers: {gj, loss, latm, bj, position, power_min, power_max}  [n, J]
    targets: {position [n, K, 2], rcs [n, K]}
    env    : {max_radar_types, rd_min, rd_max, rp_min, rp_max, episode_limit}
dB-valued entries (gt, gr, loss, latm, pn, gj) are in dB exactly as in the
reference's YAML.

Noise layout (injected uniforms in [0, 1)): ``noise[n, R*K + J]``; entry
``r*K + k`` decides the detection of target k by radar r, entry ``R*K + i``
decides whether the false target of jammer i's deception action is detected.
The reference consumes its RNG stream compactly (one draw per radar, then one
per *valid* deception action in jammer order, environment.py:341,430);
``compact_noise_for_reference`` converts between the two.
"""
from __future__ import annotations

import math

import numpy as np

# Albersheim constants exactly as evaluated by core/radar.py:67-82 with the
# defaults prfa=1e-6, m=10 (python floats, same operation order).
_PRFA = 1e-6
_M = 10
_ALB_A = float(np.log(0.62 / max(_PRFA, 1e-18)))
_ALB_ZOFF = float((5 * np.log10(_M)) / (6.2 + 4.54 / np.sqrt(_M) + 0.44))
_ALB_DEN = float(1.7 + 0.12 * _ALB_A)


def albersheim_pd(snr):
    """core/radar.py:67-82 on an array of linear SNR values (float64)."""
    snr = np.maximum(np.asarray(snr, dtype=np.float64), 0.0)
    z = snr + _ALB_ZOFF
    b = (10 * z - _ALB_A) / _ALB_DEN
    with np.errstate(over="ignore"):
        pd = 1.0 / (1.0 + np.exp(-b))
    pd = np.where(b > 700, 1.0, pd)
    pd = np.where(b < -700, 0.0, pd)
    return pd


def _db(x):
    """utils/math_utils.py:3-8 -- python float pow per element (bit-faithful)."""
    a = np.asarray(x, dtype=np.float64)
    uniq, inv = np.unique(a.ravel(), return_inverse=True)
    lin = np.array([10 ** (float(v) / 10.0) for v in uniq], dtype=np.float64)
    return lin[inv].reshape(a.shape)


def _pn_watts(pn_db):
    """core/radar.py:19  pn_watts = 10**((pn - 30) / 10)."""
    a = np.asarray(pn_db, dtype=np.float64)
    uniq, inv = np.unique(a.ravel(), return_inverse=True)
    w = np.array([10 ** ((float(v) - 30) / 10) for v in uniq], dtype=np.float64)
    return w[inv].reshape(a.shape)


class EnvOracle:
    """Batched float64 restatement of ElectromagneticEnvironment."""

    def __init__(self, spec: dict):
        rad, jam, tgt, env = spec["radars"], spec["jammers"], spec["targets"], spec["env"]
        f = lambda a: np.asarray(a, dtype=np.float64)
        self.n_envs, self.R = f(rad["pt"]).shape
        self.J = f(jam["gj"]).shape[1]
        self.K = f(tgt["rcs"]).shape[1]
        self.types = int(env["max_radar_types"])
        self.rd_min, self.rd_max = float(env["rd_min"]), float(env["rd_max"])
        self.rp_min, self.rp_max = float(env["rp_min"]), float(env["rp_max"])
        self.episode_limit = int(env["episode_limit"])
        # radar entity (core/radar.py:10-33)
        self.pt = f(rad["pt"])
        self.gt = _db(rad["gt"])
        self.gr = _db(rad["gr"])
        self.lam = f(rad["wavelength"])
        self.rloss = _db(rad["loss"])
        self.rlatm = _db(rad["latm"])
        self.pn = _pn_watts(rad["pn"])
        self.type_id = np.asarray(rad["type_id"], dtype=np.int64)
        if np.any(self.type_id < 0) or np.any(self.type_id >= self.types):
            raise ValueError("invalid type_id")  # environment.py:146-147
        self.rpos = f(rad["position"])
        self.theta_m, self.theta_a, self.t_s = f(rad["theta_m"]), f(rad["theta_a"]), f(rad["t_s"])
        self.Ga = f(rad["pulse_compression_gain"])
        self.D = f(rad["anti_jamming_factor"])
        self.threat = f(rad["threat_level"])
        # jammer entity (core/jammer.py:24-54, environment.py:184-190)
        self.gj = _db(jam["gj"])
        self.jloss = _db(jam["loss"])
        self.jlatm = _db(jam["latm"])
        self.bj = f(jam["bj"])
        self.jpos = f(jam["position"])
        self.pmin, self.pmax = f(jam["power_min"]), f(jam["power_max"])
        # protected targets
        self.tpos = f(tgt["position"])
        self.trcs = f(tgt["rcs"])
        self.state_dim = self.R * (6 + self.types) + 2 * self.J
        self.n_actions = 2 * self.R + 1
        self.step_count = np.zeros(self.n_envs, dtype=np.int64)
        self.jammer_power = np.zeros((self.n_envs, self.J))
        self.tracking = np.zeros((self.n_envs, self.R), dtype=bool)

    # ------------------------------------------------------------------ info
    def get_env_info(self):
        """environment.py:553-565"""
        return {"state_shape": self.state_dim, "obs_shape": self.state_dim,
                "n_actions": self.n_actions, "n_agents": self.J,
                "episode_limit": self.episode_limit}

    def reset(self):
        """environment.py:208-219 -- entities rebuilt, counters zeroed."""
        self.step_count[:] = 0
        self.jammer_power[:] = 0.0
        self.tracking[:] = False
        return self.get_state()

    def get_state(self):
        """environment.py:479-510 -> float32 [n, S]."""
        n, R, J, Ty = self.n_envs, self.R, self.J, self.types
        per = 6 + Ty
        s = np.zeros((n, self.state_dim), dtype=np.float64)
        for r in range(R):
            o = r * per
            s[:, o + 0] = self.pt[:, r]
            s[:, o + 1] = self.theta_m[:, r]
            s[:, o + 2] = self.t_s[:, r]
            s[np.arange(n), o + 3 + self.type_id[:, r]] = 1.0
            s[:, o + 3 + Ty] = self.theta_a[:, r]
            s[:, o + 4 + Ty] = self.rpos[:, r, 0]
            s[:, o + 5 + Ty] = self.rpos[:, r, 1]
        o = R * per
        for j in range(J):
            s[:, o + 2 * j] = self.jpos[:, j, 0]
            s[:, o + 2 * j + 1] = self.jpos[:, j, 1]
        return s.astype(np.float32)

    def get_obs(self):
        """environment.py:512-522 -> float32 [n, J, S] (state replicated per agent)."""
        s = self.get_state()
        return np.repeat(s[:, None, :], self.J, axis=1)

    def get_avail_actions(self):
        """environment.py:539-551 -> int32 ones [n, J, A]."""
        return np.ones((self.n_envs, self.J, self.n_actions), dtype=np.int32)

    # --------------------------------------------------------------- physics
    def echo_power(self):
        """core/radar.py:35-60 for every (radar, target) pair -> [n, R, K]."""
        d = self.rpos[:, :, None, :] - self.tpos[:, None, :, :]
        dist = np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1])
        dist = np.maximum(dist, 1e-6)
        num = (self.pt * self.gt * self.gr * (self.lam ** 2))[:, :, None] * self.trcs[:, None, :]
        den = ((4 * np.pi) ** 3) * (dist ** 4) * self.rloss[:, :, None] * self.rlatm[:, :, None]
        with np.errstate(divide="ignore", invalid="ignore"):
            ps = np.where(den <= 1e-18, 0.0, num / np.where(den <= 1e-18, 1.0, den))
        return ps

    def jam_gain(self):
        """Power-independent part of core/jammer.py:56-98 for every (jammer, radar)
        pair: Prj = max(0, P) * gj * gr / den, den = max(1e-9,d^2)*L*Latm*max(1e-9,Bj).
        Returns (dist [n,J,R], den [n,J,R])."""
        d = self.jpos[:, :, None, :] - self.rpos[:, None, :, :]
        dist = np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1])
        dsq = np.maximum(1e-9, dist ** 2)
        den = dsq * self.jloss[:, :, None] * self.jlatm[:, :, None] * np.maximum(1e-9, self.bj)[:, :, None]
        return dist, den

    def step(self, act_d, act_p, noise):
        """environment.py:221-477.  act_d int [n,J]; act_p float [n,J];
        noise float [n, R*K + J] (layout in the module docstring).
        Returns a dict of float64 / bool arrays."""
        n, R, J, K = self.n_envs, self.R, self.J, self.K
        act_d = np.asarray(act_d).astype(np.int64).reshape(n, J)
        p = np.clip(np.asarray(act_p, dtype=np.float64).reshape(n, J), 0.0, 1.0)
        noise = np.asarray(noise, dtype=np.float64).reshape(n, R * K + J)
        self.step_count += 1

        # --- jammer loop (environment.py:248-302)
        valid_idx = (act_d >= 1) & (act_d <= 2 * R)
        tgt = np.where(valid_idx, (act_d + 1) // 2 - 1, 0)
        jtype = act_d % 2                                   # 1 suppression, 0 deception
        power = self.pmin + p * (self.pmax - self.pmin)
        self.jammer_power = power.copy()
        prange = self.pmax - self.pmin
        norm = np.where(prange > 1e-6, (power - self.pmin) / np.where(prange > 1e-6, prange, 1.0), 0.0)
        dist, den = self.jam_gain()
        ii = np.arange(n)[:, None]
        jj = np.arange(J)[None, :]
        d_sel = dist[ii, jj, tgt]
        den_sel = den[ii, jj, tgt]
        grj = self.gr[ii, tgt]
        active = valid_idx & (power > 0) & (d_sel > 1e-6)   # an action_detail exists
        prj = np.where(den_sel <= 1e-18, 0.0,
                       (np.maximum(0.0, power) * self.gj * grj) / np.where(den_sel <= 1e-18, 1.0, den_sel))
        prj = np.maximum(0.0, prj)
        supp = active & (jtype == 1)
        dec = active & (jtype == 0)
        prjs = np.zeros((n, R))
        supp_target = np.zeros((n, R), dtype=bool)
        for j in range(J):                                  # jammer order = accumulation order
            m = supp[:, j]
            prjs[np.arange(n)[m], tgt[m, j]] += prj[m, j]
            supp_target[np.arange(n)[m], tgt[m, j]] = True

        # --- radar loop (environment.py:316-349), generalised to (r, k) pairs
        ps = self.echo_power()                               # [n,R,K]
        pn = self.pn[:, :, None]
        sig = self.Ga[:, :, None] * ps
        snr0 = np.maximum(0.0, np.where(pn > 1e-18, sig / pn, 0.0))
        den1 = (self.D * prjs)[:, :, None] + pn
        snr1_raw = np.where(den1 > 1e-18, sig / den1, 0.0)
        snr1 = np.maximum(0.0, snr1_raw)
        pd = albersheim_pd(snr1_raw)
        u = noise[:, :R * K].reshape(n, R, K)
        det = u <= pd
        tracking = det.any(axis=2)
        self.tracking = tracking

        # --- rewards (environment.py:352-457)
        pen = np.clip(-self.threat, self.rd_min, self.rd_max)
        r_d = np.where(tracking, pen, 0.0).sum(axis=1)
        r_p = (self.rp_max + (self.rp_min - self.rp_max) * norm).sum(axis=1)
        pd0 = albersheim_pd(snr0)
        red = np.maximum(0.0, pd0 - pd).sum(axis=2)          # per radar, summed over targets
        r_j_supp = np.where(supp_target, red, 0.0).sum(axis=1)
        snr_f = np.maximum(0.0, np.where(self.pn[ii, tgt] > 1e-18,
                                         (self.D[ii, tgt] * prj) / self.pn[ii, tgt], 0.0))
        pd_f = albersheim_pd(snr_f)
        uf = noise[:, R * K:]
        hit = dec & (uf <= pd_f)
        prod = np.ones((n, R))
        any_hit = np.zeros((n, R), dtype=bool)
        for j in range(J):
            m = hit[:, j]
            rows = np.arange(n)[m]
            prod[rows, tgt[m, j]] *= (1.0 - np.minimum(pd_f[m, j], 0.999999))
            any_hit[rows, tgt[m, j]] = True
        r_j_dec = np.where(any_hit, 1.0 - prod, 0.0).sum(axis=1)
        r_j = r_j_supp + r_j_dec
        reward = r_d + r_p + r_j
        terminated = self.step_count >= self.episode_limit

        # --- extensions (not in the reference)
        with np.errstate(divide="ignore", invalid="ignore"):
            jsr_db = 10.0 * np.log10((self.D * prjs)[:, :, None] / sig)
        pd_net = 1.0 - np.prod(1.0 - pd, axis=1)             # [n,K]
        return {
            "reward": reward, "r_d": r_d, "r_p": r_p, "r_j": r_j,
            "pd": pd, "detected": det, "tracking": tracking,
            "snr0": snr0, "snr1": snr1, "terminated": terminated,
            "step_count": self.step_count.copy(),
            "prj": np.where(active, prj, 0.0), "active": active, "deception": dec,
            "deception_hit": hit, "pd_false": pd_f, "norm_power": norm,
            "jsr_db": jsr_db, "pd_net": pd_net,
        }


def compact_noise_for_reference(noise_row, deception_row, R):
    """Reorder one env's slot-per-jammer noise row (K == 1) into the sequence the
    reference pops from np.random.rand: R radar draws, then one draw per valid
    deception action in jammer order (environment.py:341,430)."""
    seq = [float(x) for x in noise_row[:R]]
    for j, is_dec in enumerate(deception_row):
        if is_dec:
            seq.append(float(noise_row[R + j]))
    return seq


# --------------------------------------------------------------------------
# Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3").
# Restated from the published algorithm; the CUDA kernels use the same counter
# mapping so device-generated noise can be reproduced here.
_PHILOX_M0, _PHILOX_M1 = 0xD2511F53, 0xCD9E8D57
_PHILOX_W0, _PHILOX_W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(ctr, key):
    """ctr: 4 uint32, key: 2 uint32 (python ints) -> 4 uint32."""
    c0, c1, c2, c3 = [int(c) & 0xFFFFFFFF for c in ctr]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _PHILOX_M0 * c0
        p1 = _PHILOX_M1 * c2
        hi0, lo0 = p0 >> 32, p0 & 0xFFFFFFFF
        hi1, lo1 = p1 >> 32, p1 & 0xFFFFFFFF
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & 0xFFFFFFFF, lo1, (hi0 ^ c3 ^ k1) & 0xFFFFFFFF, lo0
        k0 = (k0 + _PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + _PHILOX_W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def u01_from_u32(x):
    """uint32 -> float32 uniform in [0, 1): top 24 bits * 2^-24 (exact in fp32)."""
    return np.float32((int(x) >> 8) * (1.0 / 16777216.0))


def device_noise(seed, stream_id, env_index, step_count, n):
    """The uniforms the CUDA kernels generate when no noise is injected:
    counter = (slot/4, step_count, env_index, stream_id), key = (seed lo, seed hi);
    value slot%4 of the block."""
    out = np.empty(n, dtype=np.float32)
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    for blk in range((n + 3) // 4):
        r = philox4x32_10((blk, step_count, env_index, stream_id), key)
        for q in range(4):
            if blk * 4 + q < n:
                out[blk * 4 + q] = u01_from_u32(r[q])
    return out
