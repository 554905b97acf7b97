"""Batched electromagnetic environment on the GPU behind the reference's class API.

Drop-in for ``simulation/environment.py:29-573`` of the reference
(``ElectromagneticEnvironment``): same constructor, ``reset / step / get_state /
get_obs / get_agent_obs / get_avail_actions / get_env_info / close``.

Two modes
  * **single-instance shim** (default, what runners/episode_runner.py and main.py use):
    one episode, exact reference return types -- ``step`` takes a list of
    ``(T_i, P_i)`` tuples and returns ``(list[np.float32[S]], float, bool, info)`` with
    the info keys of environment.py:463-473.
  * **batched** (``n_envs=`` given): the same methods with a leading ``n_envs`` axis on
    CUDA tensors: ``step((act_d int32[n,J], act_p float32[n,J]), noise=None)`` returns
    ``(obs f32[n,J,S], reward f32[n], terminated bool[n], info)``; thousands of episodes
    advance in one fused kernel launch (csrc/env_step.cuh).

All arithmetic happens in the CUDA library (include/macjd.h: macjd_env_step /
macjd_env_reset); there is no CPU path.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _native as N
from .scenario import ScenarioTables, load_sim_config, spec_from_config

DEFAULT_SIM_CONFIG_PATH = "config/simulation_config.yaml"

# Albersheim constants with the reference's defaults prfa=1e-6, m=10, evaluated the way
# core/radar.py:67-82 evaluates them (float64).
_ALB_A = float(np.log(0.62 / max(1e-6, 1e-18)))
_ALB_ZOFF = float((5 * np.log10(10)) / (6.2 + 4.54 / np.sqrt(10) + 0.44))
_ALB_DEN = float(1.7 + 0.12 * _ALB_A)


class ElectromagneticEnvironment:
    # every agent's observation IS the global state (environment.py:512-522: get_obs returns get_state once per
    # jammer); a replay ring may therefore keep the state only (utils/replay_buffer.py: shared_obs)
    obs_is_replicated_state = True

    def __init__(self, config, sim_config_path=DEFAULT_SIM_CONFIG_PATH, *, n_envs=None, spec=None,
                 device=None, seed=None, auto_reset=False, share_scenario=False, derive_tables=True, _lib=None):
        """config: RL config namespace (reads num_jammers / num_radars / episode_limit with
        the reference's fallbacks, environment.py:82-84).  ``spec`` (a raw scenario spec,
        see scenario.py) replaces the YAML for heterogeneous / synthetic scenarios."""
        self._lib = _lib if _lib is not None else N.get_lib()
        self.batched = n_envs is not None or spec is not None
        if spec is None:
            self.sim_config = load_sim_config(sim_config_path)
            spec = spec_from_config(
                self.sim_config, n_envs=int(n_envs or 1),
                num_radars=getattr(config, "num_radars", None),
                num_jammers=getattr(config, "num_jammers", None),
                episode_limit=getattr(config, "episode_limit", 100))
        elif hasattr(config, "episode_limit"):
            spec = dict(spec, env=dict(spec["env"], episode_limit=int(config.episode_limit)))
        tabs = ScenarioTables(spec)
        self.tables = tabs
        self.n_envs = tabs.n_envs
        self.num_jammers, self.num_radars, self.num_targets = tabs.J, tabs.R, tabs.K
        self.max_radar_types = tabs.types
        self.episode_limit = tabs.episode_limit
        self.rd_min_penalty, self.rd_max_penalty = tabs.rd_min, tabs.rd_max
        self.rp_min_penalty, self.rp_max_penalty = tabs.rp_min, tabs.rp_max
        self.action_dim_discrete = tabs.n_actions
        self.action_dim_continuous = 1
        self.state_dim = self.obs_dim = self.agent_obs_dim = tabs.state_dim
        self.seed = int(seed if seed is not None else getattr(config, "seed", 0) or 0)
        self.auto_reset = bool(auto_reset)

        if device is None:
            device = "cuda"
        self.device = torch.device(device)
        if self.device.type != "cuda" and _lib is None:
            raise N.MacjdError("ElectromagneticEnvironment runs on CUDA only (no CPU fallback)")
        dev = self.device
        n, J, R, K, S, A = self.n_envs, tabs.J, tabs.R, tabs.K, tabs.state_dim, tabs.n_actions
        if share_scenario:
            # one scenario for every env: [row] table, env_stride 0 (5x less table traffic)
            self._tab_dev = torch.from_numpy(np.ascontiguousarray(tabs.data[:, 0])).to(dev)
            row_stride, env_stride = 1, 0
        else:
            self._tab_dev = torch.from_numpy(tabs.data).to(dev)
            row_stride, env_stride = n, 1
        self._ctab = N.EnvTables(
            n_envs=n, n_jammers=J, n_radars=R, n_targets=K, n_types=tabs.types,
            episode_limit=tabs.episode_limit, data=self._tab_dev.data_ptr(),
            row_stride=row_stride, env_stride=env_stride,
            rd_min=tabs.rd_min, rd_max=tabs.rd_max, rp_min=tabs.rp_min, rp_max=tabs.rp_max,
            alb_a=_ALB_A, alb_zoff=_ALB_ZOFF, alb_den=_ALB_DEN, derived=None)
        # Scenario-only terms of the step (echo power and no-jamming Pd per radar-target pair, link denominators,
        # state rows): derived once here, read by every step (csrc/env_step2.cuh).  derive_tables=False keeps the
        # kernel that works from the raw tables on every step (round 1's; same results).
        if derive_tables and R <= 32 and J <= 32:       # (the group-per-env kernel covers up to 32 jammers / radars)
            nbytes = int(self._lib.lib.macjd_env_derived_bytes(N.C.byref(self._ctab)))
            self._derived_dev = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
            self._lib.callv("macjd_env_prepare", self._ctx_for(dev), self._ctab, self._derived_dev)
            self._ctab.derived = self._derived_dev.data_ptr()

        f32 = dict(dtype=torch.float32, device=dev)
        u8 = dict(dtype=torch.uint8, device=dev)
        self.step_count = torch.zeros(n, dtype=torch.int32, device=dev)
        self.reward = torch.zeros(n, **f32)
        self.reward64 = torch.zeros(n, dtype=torch.float64, device=dev)
        self.r_d, self.r_p, self.r_j = (torch.zeros(n, **f32) for _ in range(3))
        self.terminated = torch.zeros(n, **u8)
        self.pd = torch.zeros(R * K, n, **f32)
        self.detected = torch.zeros(R * K, n, **u8)
        self.tracking = torch.zeros(R, n, **u8)
        self.snr0 = torch.zeros(R * K, n, **f32)
        self.snr1 = torch.zeros(R * K, n, **f32)
        self.jsr_db = torch.zeros(R * K, n, **f32)
        self.pd_net = torch.zeros(K, n, **f32)
        self.jam_power = torch.zeros(J, n, **f32)
        self.state = torch.zeros(n, S, **f32)
        self.obs = torch.zeros(n, J, S, **f32)
        self.avail = torch.zeros(n, J, A, **u8)
        self._last_actions = np.zeros((J, 2))
        self._reset_device()

        if not self.batched:
            # same one-time summary the reference prints (environment.py:114-121)
            print(f"Environment Initialized: {J} Jammers, {R} Radars (from {sim_config_path})")
            print(f"State Dimension: {S}")
            print(f"Action Dimension (Discrete): {A}")
            print(f"Episode Limit: {self.episode_limit}")

    # ------------------------------------------------------------------ native calls
    def _ctx(self):
        return self._ctx_for(self.device)

    @staticmethod
    def _ctx_for(device):
        if device.type == "cuda":
            return N.torch_ctx(device)
        return N.Ctx(device=0, reserved=0, stream=None)

    def _io(self, act_d=None, act_p=None, noise=None, out=None):
        p = N.ptr
        io = N.EnvIO(
            act_d=p(act_d), act_p=p(act_p), noise=p(noise), seed=self.seed,
            auto_reset=int(self.auto_reset), flags=0, step_count=p(self.step_count),
            reward=p(self.reward), r_d=p(self.r_d), r_p=p(self.r_p), r_j=p(self.r_j),
            reward64=p(self.reward64), terminated=p(self.terminated),
            pd=p(self.pd), detected=p(self.detected), tracking=p(self.tracking),
            snr0=p(self.snr0), snr1=p(self.snr1), jsr_db=p(self.jsr_db), pd_net=p(self.pd_net),
            jam_power=p(self.jam_power), state=p(self.state), obs=p(self.obs), avail=p(self.avail))
        if out:
            # redirect selected outputs (e.g. into the slices of a rollout trajectory);
            # a value of None drops an optional output
            for k, v in out.items():
                setattr(io, k, p(v))
        return io

    def _reset_device(self, out=None):
        self._lib.call("macjd_env_reset", self._ctx(), self._ctab, self._io(out=out))

    def step_device(self, act_d, act_p, noise=None, out=None):
        """Enqueue one step for all envs.  act_d int32 [n,J], act_p float32 [n,J] and
        optional noise float32 [n, R*K+J] must be contiguous tensors on the env device.
        Results land in the persistent output buffers (self.reward, self.obs, ...) unless
        ``out`` redirects them ({EnvIO field: tensor})."""
        self._lib.call("macjd_env_step", self._ctx(), self._ctab, self._io(act_d, act_p, noise, out))

    def host_buffers(self, pinned=True):
        """Host arrays for ``step_host`` (page-locked unless ``pinned=False``): the discrete / continuous
        actions going in and reward, terminated, next observations coming out, as CPU tensors."""
        n, J, S = self.n_envs, self.num_jammers, self.state_dim
        pin = bool(pinned) and self.device.type == "cuda"
        mk = lambda shape, dt: torch.zeros(shape, dtype=dt, pin_memory=pin)
        return {"act_d": mk((n, J), torch.int32), "act_p": mk((n, J), torch.float32), "reward": mk((n,), torch.float32),
                "terminated": mk((n,), torch.uint8), "obs": mk((n, J, S), torch.float32)}

    def step_host(self, host):
        """environment.py:221-477 for callers that live on the host, as the reference's runner does:
        ``host`` maps act_d / act_p (in) and any of reward / terminated / obs / state (out) to contiguous
        CPU tensors or numpy arrays (see ``host_buffers``).  One C call: copies in, fused step, copies
        out, stream drained on return (include/macjd.h: macjd_env_step_host)."""
        self._lib.call("macjd_env_step_host", self._ctx(), self._ctab, *self.host_step_args(host))

    def host_step_args(self, host):
        """The (io, host) structs of one ``step_host`` call, cached on the caller's buffers."""
        key = (tuple(host), tuple([v.data_ptr() if hasattr(v, "data_ptr") else N.ptr(v) for v in host.values()]))
        c = getattr(self, "_host_cache", None)
        if c is None or c["key"] != key:
            n, J = self.n_envs, self.num_jammers
            if not hasattr(self, "_act_d_dev"):
                self._act_d_dev = torch.zeros(n, J, dtype=torch.int32, device=self.device)
                self._act_p_dev = torch.zeros(n, J, dtype=torch.float32, device=self.device)
            unknown = set(host) - {"act_d", "act_p", "reward", "terminated", "obs", "state"}
            if unknown or "act_d" not in host or "act_p" not in host:
                raise ValueError(f"step_host: need act_d and act_p, unknown keys {sorted(unknown)}")
            pinned = all(torch.is_tensor(v) and v.is_pinned() for v in host.values()) and self.device.type == "cuda"
            hs = N.EnvHost(flags=N.HOST_PINNED if pinned else 0, **{k: N.ptr(v) for k, v in host.items()})
            # a caller that keeps the state only (one observation row per env) does not get the per-jammer copies written
            io = self._io(self._act_d_dev, self._act_p_dev, out=None if "obs" in host else {"obs": None})
            c = self._host_cache = {"key": key, "host": hs, "io": io, "keep": dict(host)}
        return c["io"], c["host"]

    # ------------------------------------------------------------------ reference API
    def reset(self):
        """environment.py:208-219."""
        self._reset_device()
        return self.get_state()

    def step(self, actions, noise=None):
        """environment.py:221-477.  See the module docstring for the two calling modes."""
        J = self.num_jammers
        if self.batched:
            act_d, act_p = actions
            act_d = torch.as_tensor(act_d, device=self.device).to(torch.int32).reshape(self.n_envs, J).contiguous()
            act_p = torch.as_tensor(act_p, device=self.device).to(torch.float32).reshape(self.n_envs, J).contiguous()
            if noise is not None:
                noise = torch.as_tensor(noise, device=self.device).to(torch.float32).contiguous()
            self.step_device(act_d, act_p, noise)
            info = {"radar_pds": self.pd.view(self.num_radars, self.num_targets, -1).permute(2, 0, 1),
                    "radar_tracking": self.tracking.t().bool(),
                    "detected": self.detected.view(self.num_radars, self.num_targets, -1).permute(2, 0, 1).bool(),
                    "snr_no_jamming": self.snr0.view(self.num_radars, self.num_targets, -1).permute(2, 0, 1),
                    "snr_with_jamming": self.snr1.view(self.num_radars, self.num_targets, -1).permute(2, 0, 1),
                    "jsr_db": self.jsr_db.view(self.num_radars, self.num_targets, -1).permute(2, 0, 1),
                    "pd_networked": self.pd_net.t(),
                    "r_d": self.r_d, "r_p": self.r_p, "r_j": self.r_j,
                    "jammer_power": self.jam_power.t(), "step_count": self.step_count}
            return self.obs, self.reward, self.terminated.bool(), info

        if len(actions) != J:
            raise ValueError(f"Received {len(actions)} actions, but expected {J}")
        act_d_h = np.array([int(a[0]) for a in actions], dtype=np.int32).reshape(1, J)
        act_p_h = np.array([float(a[1]) for a in actions], dtype=np.float32).reshape(1, J)
        for i in range(J):
            if act_d_h[0, i] > 2 * self.num_radars:
                print(f"Warning: Jammer {i} chose invalid discrete action T_i={int(act_d_h[0, i])}")
        noise_t = None
        if noise is not None:
            noise_t = torch.as_tensor(np.asarray(noise, dtype=np.float32).reshape(1, -1), device=self.device)
        self.step_device(torch.from_numpy(act_d_h).to(self.device), torch.from_numpy(act_p_h).to(self.device), noise_t)
        self._last_actions = np.stack([act_d_h[0].astype(np.float64), np.clip(act_p_h[0], 0, 1).astype(np.float64)], axis=1)
        pd = self.pd[:, 0].double().cpu().numpy()
        tracking = self.tracking[:, 0].cpu().numpy().astype(bool)
        threat = self.tables.data[[16 * r + 9 for r in range(self.num_radars)], 0]
        info = {
            "radar_pds": pd,
            "radar_states": [{"state": "TRACK" if tracking[r] else "SEARCH", "threat": float(threat[r]),
                              "is_tracking": bool(tracking[r]), "locked_target": None} for r in range(self.num_radars)],
            "snr_no_jamming": self.snr0[:, 0].double().cpu().numpy(),
            "snr_with_jamming": self.snr1[:, 0].double().cpu().numpy(),
            "r_d": float(self.r_d[0]), "r_p": float(self.r_p[0]), "r_j": float(self.r_j[0]),
            "jammer_actions": self._jammer_action_details(act_d_h[0]),
        }
        reward = np.float64(self.reward64[0].item())
        return self.get_obs(), reward, bool(self.terminated[0].item()), info

    def _jammer_action_details(self, act_d):
        """The `jammer_actions` debugging list of environment.py:288-295 (shim mode only):
        one entry per action that actually radiated at a radar (valid index, power > 0,
        distance > 1e-6), with the received jamming power of core/jammer.py:73-98 (host arithmetic
        on the float64 scenario tables; not on the hot path)."""
        out = []
        tab = self.tables.data[:, 0]
        R, J = self.num_radars, self.num_jammers
        power = self.jam_power[:, 0].double().cpu().numpy()
        for i, T in enumerate(act_d):
            if not (1 <= T <= 2 * R and power[i] > 0):
                continue
            tgt = int((T + 1) // 2 - 1)
            jr, rr = 16 * R + 8 * i, 16 * tgt
            dist = float(np.hypot(tab[jr + 4] - tab[rr + 10], tab[jr + 5] - tab[rr + 11]))
            if dist <= 1e-6:
                continue
            den = max(1e-9, dist ** 2) * tab[jr + 1] * tab[jr + 2] * max(1e-9, tab[jr + 3])
            prj = 0.0 if den <= 1e-18 else max(0.0, max(0.0, power[i]) * tab[jr + 0] * tab[rr + 2] / den)
            out.append({"jammer_idx": i, "target_idx": tgt, "type": int(T % 2), "power": float(power[i]),
                        "received_power": float(prj)})
        return out

    def get_state(self):
        """environment.py:479-510."""
        if self.batched:
            return self.state
        return self.state[0].cpu().numpy()

    def get_obs(self):
        """environment.py:512-522."""
        if self.batched:
            return self.obs
        s = self.state[0].cpu().numpy()
        return [s for _ in range(self.num_jammers)]

    def get_agent_obs(self, agent_id):
        """environment.py:524-537."""
        if not (0 <= agent_id < self.num_jammers):
            raise ValueError(f"Invalid agent_id {agent_id} for {self.num_jammers} jammers.")
        if self.batched:
            return self.obs[:, agent_id]
        return self.get_state()

    def get_avail_actions(self):
        """environment.py:539-551."""
        if self.batched:
            return self.avail
        a = self.avail[0].cpu().numpy().astype(np.int32)
        return [a[j] for j in range(self.num_jammers)]

    def get_env_info(self):
        """environment.py:553-565."""
        return {"state_shape": self.state_dim, "obs_shape": self.state_dim,
                "n_actions": self.action_dim_discrete, "n_agents": self.num_jammers,
                "episode_limit": self.episode_limit}

    def close(self):
        """environment.py:567-573."""
        if not self.batched:
            print("Closing Electromagnetic Environment.")
