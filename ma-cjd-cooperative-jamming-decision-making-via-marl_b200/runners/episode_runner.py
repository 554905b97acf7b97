"""Episode rollouts (drop-in for runners/episode_runner.py).

``EpisodeRunner`` follows the reference protocol one environment at a time (it works with
the single-instance shim of the environment and stores one host-staged episode per call,
episode_runner.py:27-181).  ``BatchedEpisodeRunner`` is the device-resident version: all
``n_envs`` episodes advance together, the agent and step kernels write straight into
time-major trajectory buffers in HBM and the finished episodes go to the replay ring with
one copy kernel -- no host round trip inside an episode.
"""
from __future__ import annotations

from functools import partial

import numpy as np
import torch


class EpisodeBatch:
    """Host staging of one episode (episode_runner.py:187-277): zero-initialised arrays with
    T (+1) slots; `update_last` only writes if the episode ended early (reference fact 9)."""

    T_PLUS_1 = ("state", "obs", "avail_actions", "hidden_state")

    def __init__(self, args, max_seq_length, n_agents, obs_shape):
        self.args = args
        self.max_seq_length = max_seq_length
        self.n_agents = n_agents
        self.obs_shape = obs_shape
        self.scheme = self._get_scheme(args)
        self.data = {}
        self.t = 0

    def _get_scheme(self, args):
        info = args.env_info
        return {"state": ((info["state_shape"],), None, np.float32),
                "obs": ((self.obs_shape,), "agents", np.float32),
                "actions_discrete": ((1,), "agents", np.int32),
                "actions_continuous": ((1,), "agents", np.float32),
                "avail_actions": ((info["n_actions"],), "agents", np.int64),
                "reward": ((1,), None, np.float32),
                "terminated": ((1,), None, np.bool_),
                "hidden_state": ((args.rnn_hidden_dim,), "agents", np.float32)}

    def _init_data(self):
        self.data = {}
        for key, (shape, group, dtype) in self.scheme.items():
            n_t = self.max_seq_length + 1 if key in self.T_PLUS_1 else self.max_seq_length
            full = (n_t, self.n_agents) + shape if group == "agents" else (n_t,) + shape
            self.data[key] = np.zeros(full, dtype=dtype)
        self.t = 0

    def push(self, transition_data):
        if self.t == 0:
            self._init_data()
        if self.t < self.max_seq_length:
            for key, value in transition_data.items():
                if key in self.data:
                    self.data[key][self.t] = value
            self.t += 1
        else:
            print("Warning: Episode length exceeded max_seq_length. Data not stored.")

    def update_last(self, last_data):
        if self.t < self.max_seq_length:
            for key, value in last_data.items():
                if key in self.data:
                    self.data[key][self.t] = value

    def get_batch_data(self):
        out = {}
        for key in self.scheme:
            n = self.t + 1 if key in self.T_PLUS_1 else self.t
            out[key] = [self.data[key][:n]]
        return out


class EpisodeRunner:
    """One episode per ``run`` through the reference-facing API (episode_runner.py:5-184)."""

    def __init__(self, env, mac, buffer, args):
        self.env, self.mac, self.buffer, self.args = env, mac, buffer, args
        info = self.env.get_env_info()
        self.episode_limit = info["episode_limit"]
        self.n_agents = info["n_agents"]
        self.t = 0
        self.t_env = 0
        self.new_batch = partial(EpisodeBatch, self.args, self.episode_limit, self.n_agents, info["obs_shape"])
        self.device = mac.device if hasattr(mac, "device") else torch.device("cpu")

    def run(self, test_mode=False):
        batch = self.new_batch()
        self.mac.init_hidden(batch_size=1)
        self.env.reset()
        state = self.env.get_state()
        obs_list = self.env.get_obs()
        terminated, step, episode_return = False, 0, 0
        log = {"r": [], "r_d": [], "r_p": [], "r_j": [], "a_d": [], "a_c": []}
        info = {}
        while not terminated:
            avail_list = self.env.get_avail_actions()
            obs_t = torch.as_tensor(np.array(obs_list), dtype=torch.float32).unsqueeze(0).to(self.device)
            avail_t = torch.as_tensor(np.array(avail_list), dtype=torch.long).unsqueeze(0).to(self.device)
            a_t, p_t = self.mac.select_actions(obs_t, avail_t, self.t_env, test_mode=test_mode)
            hidden = self.mac.hidden_states.detach().cpu().reshape(self.n_agents, -1).numpy()
            a_d = a_t.detach().squeeze(0).cpu().numpy()
            a_c = p_t.detach().squeeze(0).cpu().numpy()
            next_obs, reward, terminated, info = self.env.step([(d[0], c[0]) for d, c in zip(a_d, a_c)])
            log["r"].append(reward); log["r_d"].append(info.get("r_d", 0)); log["r_p"].append(info.get("r_p", 0))
            log["r_j"].append(info.get("r_j", 0)); log["a_d"].append(a_d); log["a_c"].append(a_c)
            episode_return += reward
            batch.push({"state": np.array(state), "obs": np.array(obs_list), "actions_discrete": a_d,
                        "actions_continuous": a_c, "avail_actions": np.array(avail_list),
                        "reward": np.array([reward]), "terminated": np.array([terminated]), "hidden_state": hidden})
            state = self.env.get_state()
            obs_list = next_obs
            step += 1
            self.t_env += 1
            if terminated or step >= self.episode_limit:
                batch.update_last({"state": np.array(self.env.get_state()), "obs": np.array(next_obs),
                                   "avail_actions": np.array(self.env.get_avail_actions())})
                break
        if not test_mode:
            self.buffer.store_episode(batch.get_batch_data())
        a_d_all = np.array(log["a_d"]) if log["a_d"] else np.empty((0, self.n_agents, 1))
        a_c_all = np.array(log["a_c"]) if log["a_c"] else np.empty((0, self.n_agents, 1))
        counts = np.bincount(a_d_all.flatten().astype(int), minlength=self.args.n_actions)[:self.args.n_actions] \
            if a_d_all.size else np.zeros(self.args.n_actions)
        run_info = {"episode_length": step, "episode_return": episode_return,
                    "avg_step_reward": np.mean(log["r"]) if log["r"] else 0,
                    "avg_r_d": np.mean(log["r_d"]) if log["r_d"] else 0,
                    "avg_r_p": np.mean(log["r_p"]) if log["r_p"] else 0,
                    "avg_r_j": np.mean(log["r_j"]) if log["r_j"] else 0,
                    "avg_power_overall": float(np.mean(a_c_all)) if a_c_all.size else 0.0,
                    "action_distribution": counts / max(1, counts.sum())}
        if "individual_rewards" in info:
            run_info["individual_rewards_final"] = info["individual_rewards"]
        return run_info

    def close_env(self):
        self.env.close()


class BatchedEpisodeRunner:
    """All ``n_envs`` episodes of a batched environment advance together on the device.

    Per timestep: one fused agent kernel (reads obs / avail of slot t, writes actions, powers
    and h_t into slot t) and one fused env-step kernel (reads them, writes reward / terminated
    into slot t and state / obs / avail into slot t+1).  ``run`` returns the same ``run_info``
    keys as the reference runner, averaged over the batch, and stores the episodes in the
    replay ring (``store_rollout``)."""

    def __init__(self, env, mac, buffer, args):
        assert env.batched, "BatchedEpisodeRunner needs ElectromagneticEnvironment(n_envs=...)"
        self.env, self.mac, self.buffer, self.args = env, mac, buffer, args
        info = env.get_env_info()
        self.episode_limit, self.n_agents = info["episode_limit"], info["n_agents"]
        self.n_envs = env.n_envs
        # total single-environment steps, the reference's meaning of t_env (episode_runner.py:117: one tick per
        # env.step of ONE env; mac.py:96 anneals epsilon on it): a batched timestep advances it by n_envs, so
        # an exploration schedule written for the reference decays over the same amount of experience
        self.t_env = 0
        dev = env.device
        n, T, Nn, S, A, H = self.n_envs, self.episode_limit, self.n_agents, info["state_shape"], info["n_actions"], args.rnn_hidden_dim
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)
        self.traj = {"state": z((T + 1, n, S), torch.float32), "obs": z((T + 1, n, Nn, S), torch.float32),
                     "actions_discrete": z((T, n, Nn, 1), torch.int32), "actions_continuous": z((T, n, Nn, 1), torch.float32),
                     "avail_actions": z((T + 1, n, Nn, A), torch.uint8), "reward": z((T, n, 1), torch.float32),
                     "terminated": z((T, n, 1), torch.uint8), "hidden_state": z((T + 1, n, Nn, H), torch.float32)}
        self.r_parts = z((3, T, n), torch.float32)
        # True: a timestep is one macjd_rollout_step call (the library fuses agent step and env step into one launch
        # when it can, include/macjd.h); False: macjd_agent_forward + macjd_env_step.  Same results.
        self.fused_step = bool(getattr(args, "fused_rollout_step", True))
        # run(): the whole episode as one macjd_rollout_steps call (one launch when fused) instead of one call per step
        self.whole_episode_launch = bool(getattr(args, "whole_episode_launch", True))

    def _build_step_structs(self):
        """All pointers of a timestep are fixed (persistent trajectory buffers), so the C structs
        of its two launches are built once; per step only epsilon / rng_step / test_mode change.
        Keeps the host cost of a step at two ctypes calls (the rollout stays GPU-bound)."""
        from .. import _native as N
        tr, mac, env = self.traj, self.mac, self.env
        n, Nn = self.n_envs, self.n_agents
        p = N.ptr
        self._agent_io, self._env_io = [], []
        T = self.episode_limit
        for t in range(T):
            # The recurrent state is chained through the trajectory's own h_t records: step t reads slot t - 1
            # (step 0 starts from zeros, reset()) and writes slot t; only the last step also updates
            # mac.hidden_states.  One 4 MB store per step instead of two: 41.7 -> 39.9 us per flushed step
            # at 4 096 envs (tools/step_hidden_seq.py).  Steps of one episode must therefore all take this path
            # (or all the injected-draws path below), in order -- which is how run() and the graph use them.
            self._agent_io.append(N.AgentIO(
                n_rows=n * Nn, n_steps=1, obs=p(tr["obs"][t]), hidden=p(mac.hidden_states) if t == T - 1 else None,
                hidden_in=p(tr["hidden_state"][t - 1]) if t > 0 else None, hidden_zero_init=1 if t == 0 else 0,
                test_mode=0, tile_rows=0, path=mac.agent.path, hidden_seq=p(tr["hidden_state"][t]),
                avail=p(tr["avail_actions"][t]), epsilon=0.0, rng_step=0, seed=mac.seed & 0xFFFFFFFFFFFFFFFF,
                actions=p(tr["actions_discrete"][t]), power=p(tr["actions_continuous"][t])))
            self._env_io.append(env._io(tr["actions_discrete"][t], tr["actions_continuous"][t], None, out={
                "reward": tr["reward"][t], "terminated": tr["terminated"][t], "r_d": self.r_parts[0, t],
                "r_p": self.r_parts[1, t], "r_j": self.r_parts[2, t], "state": tr["state"][t + 1],
                "obs": tr["obs"][t + 1], "avail": tr["avail_actions"][t + 1]}))
            self._env_io[-1].flags = N.ENV_FOLLOWS_AGENT      # step() launches it right behind the agent kernel
        self._structs_for = (mac.hidden_states.data_ptr(), mac.agent.path)

    def step(self, t, test_mode=False, noise=None, u_eps=None, rand_actions=None):
        """Timestep t of the current episodes: agent act + env step, both into the trajectory."""
        tr, mac, env = self.traj, self.mac, self.env
        n, Nn = self.n_envs, self.n_agents
        eps = mac.action_selector.anneal(self.t_env, test_mode)
        mac._rng_step += 1
        if noise is None and u_eps is None and rand_actions is None:
            # fast path: cached structs
            if getattr(self, "_structs_for", None) != (mac.hidden_states.data_ptr(), mac.agent.path):
                self._build_step_structs()
            aio = self._agent_io[t]
            aio.epsilon, aio.rng_step, aio.test_mode = float(eps), mac._rng_step & 0xFFFFFFFF, int(test_mode)
            lib, ctx = mac.agent.lib(), mac.agent._ctx()
            if self.fused_step:
                # one call; ONE launch when the CTA-pair kernel can run the env step of its rows' envs itself
                lib.call("macjd_rollout_step", ctx, mac.agent.packed().cstruct(), aio, env._ctab, self._env_io[t])
            else:
                lib.call("macjd_agent_forward", ctx, mac.agent.packed().cstruct(), aio)
                lib.call("macjd_env_step", ctx, env._ctab, self._env_io[t])
            self.t_env += self.n_envs
            return
        mac.agent.run(tr["obs"][t].view(1, n * Nn, -1), mac.hidden_states, avail=tr["avail_actions"][t],
                      epsilon=eps, test_mode=test_mode, u_eps=u_eps, rand_actions=rand_actions, seed=mac.seed,
                      rng_step=mac._rng_step, select=True, want_hidden_seq=True,
                      out={"actions": tr["actions_discrete"][t], "power": tr["actions_continuous"][t],
                           "hidden_seq": tr["hidden_state"][t]})
        env.step_device(tr["actions_discrete"][t], tr["actions_continuous"][t], noise,
                        out={"reward": tr["reward"][t], "terminated": tr["terminated"][t],
                             "r_d": self.r_parts[0, t], "r_p": self.r_parts[1, t], "r_j": self.r_parts[2, t],
                             "state": tr["state"][t + 1], "obs": tr["obs"][t + 1], "avail": tr["avail_actions"][t + 1]})
        self.t_env += self.n_envs

    def rollout(self, t0, n_steps, test_mode=False):
        """Timesteps t0 .. t0 + n_steps - 1 of the current episodes as ONE library call (include/macjd.h:
        macjd_rollout_steps) -- one launch when the CTA-pair kernel can run its rows' env steps itself: the recurrent
        state then stays in shared memory across the steps and there is no per-step launch, prologue or hand-over.
        Same trajectory as ``step(t)`` for each t (the exploration schedule is evaluated per step on the host and
        handed over as an array; the Philox counters advance by one per step)."""
        from .. import _native as N
        tr, mac, env = self.traj, self.mac, self.env
        n, Nn, T = self.n_envs, self.n_agents, self.episode_limit
        assert 0 <= t0 and n_steps >= 1 and t0 + n_steps <= T
        if getattr(self, "_structs_for", None) != (mac.hidden_states.data_ptr(), mac.agent.path):
            self._build_step_structs()
        if self.__dict__.get("_eps_dev") is None or self._eps_dev.device != env.device:
            self._eps_dev = torch.zeros(T, dtype=torch.float32, device=env.device)
            self._eps_host = torch.zeros(T, dtype=torch.float32, pin_memory=env.device.type == "cuda")
            self._multi_io = {}
        for i in range(n_steps):
            self._eps_host[t0 + i] = mac.action_selector.anneal(self.t_env + i * n, test_mode)
        self._eps_dev[t0:t0 + n_steps].copy_(self._eps_host[t0:t0 + n_steps], non_blocking=True)
        key = (t0, n_steps, mac.hidden_states.data_ptr(), mac.agent.path)
        aio = self._multi_io.get(key)
        if aio is None:
            p = N.ptr
            last = t0 + n_steps == T
            aio = self._multi_io[key] = N.AgentIO(
                n_rows=n * Nn, n_steps=n_steps, obs=p(tr["obs"][t0]), hidden=p(mac.hidden_states) if last else None,
                hidden_in=p(tr["hidden_state"][t0 - 1]) if t0 > 0 else None, hidden_zero_init=1 if t0 == 0 else 0,
                test_mode=0, tile_rows=0, path=mac.agent.path, hidden_seq=p(tr["hidden_state"][t0]),
                avail=p(tr["avail_actions"][t0]), epsilon=0.0, rng_step=0, seed=mac.seed & 0xFFFFFFFFFFFFFFFF,
                actions=p(tr["actions_discrete"][t0]), power=p(tr["actions_continuous"][t0]),
                epsilon_dev=self._eps_dev.data_ptr() + 4 * t0)
        aio.rng_step, aio.test_mode = (mac._rng_step + 1) & 0xFFFFFFFF, int(test_mode)
        mac.agent.lib().call("macjd_rollout_steps", mac.agent._ctx(), mac.agent.packed().cstruct(), aio, env._ctab, self._env_io[t0])
        mac._rng_step += n_steps
        self.t_env += n_steps * n

    def step_host(self, obs, avail, host, test_mode=False):
        """One iteration of the reference's loop (episode_runner.py:119-165: select_actions, env.step) for a
        caller whose buffers live on the HOST: ``obs`` [n, N, obs] float32 and ``avail`` [n, N, A] uint8 in;
        ``host`` as for ``ElectromagneticEnvironment.step_host`` -- its act_d / act_p receive the chosen
        actions, reward / terminated / obs / state the env outputs.  One C call and one stream drain
        (include/macjd.h: macjd_rollout_step_host); the actions reach the env kernel without leaving the
        device.  Same results as ``mac.select_actions_host`` followed by ``env.step_host``."""
        mac, env = self.mac, self.env
        # steady state (the caller hands in the same objects every step): identity checks, then one C call
        c = self.__dict__.get("_host_step_cache")
        if (c is not None and c[0] is obs and c[1] is avail and c[2] is host and c[3] is mac.hidden_states
                and c[4] == mac.agent.path and c[5] == tuple(map(id, host.values()))):
            aio, ahost, eio, ehost = c[6]
            eps = mac.action_selector.anneal(self.t_env, test_mode)
            mac._rng_step += 1
            aio.epsilon, aio.rng_step, aio.test_mode = float(eps), mac._rng_step & 0xFFFFFFFF, int(test_mode)
            w = mac.agent.packed().cstruct()
        else:
            w, aio, ahost, _, _ = mac.host_step_args(obs, avail, self.t_env, test_mode, actions_out=host["act_d"],
                                                     power_out=host["act_p"])
            eio, ehost = env.host_step_args(host)
            # (the controller's and the env's own caches keep every buffer alive, so the ids stay unique)
            self._host_step_cache = (obs, avail, host, mac.hidden_states, mac.agent.path, tuple(map(id, host.values())),
                                     (aio, ahost, eio, ehost))
        mac.agent.lib().call("macjd_rollout_step_host", mac.agent._ctx(), w, aio, ahost, env._ctab, eio, ehost)
        self.t_env += env.n_envs

    def reset(self):
        tr, mac = self.traj, self.mac
        hs = mac.hidden_states
        if hs is not None and hs.shape[0] == self.n_envs * self.n_agents and hs.device == self.env.device:
            hs.zero_()                   # same buffer: launch structs and the episode graph stay valid
        else:
            mac.init_hidden(batch_size=self.n_envs)
        self.env._reset_device(out={"state": tr["state"][0], "obs": tr["obs"][0], "avail": tr["avail_actions"][0]})

    # ------------------------------------------------------------------ whole episode as one CUDA graph
    def _episode_graph(self, test_mode):
        """The 2 x episode_limit launches of an episode captured once and replayed (opt-in: run(use_graph=True)).
        Kernel parameters are frozen in a graph, so the two per-step scalars that change between episodes
        -- epsilon and the Philox step counter -- are read from device memory (macjd_agent_io.epsilon_dev /
        rng_step_dev).  Rebuilt when any captured address changes.  Measured on B200 at 4 096 envs: with the
        cached launch structs a step costs the host 19 us against 40 us of GPU time, so the per-step launches
        are already GPU-bound (43.2 us per step for a whole run()) and the replay is not faster (46.3 us);
        it pays off only when the host is slow or shared."""
        from .. import _native as N
        mac, env, tr = self.mac, self.env, self.traj
        pk = mac.agent.packed()
        key = (bool(test_mode), mac.hidden_states.data_ptr(), pk.buffer.data_ptr(),
               pk.tc_buffer.data_ptr() if pk.tc_buffer is not None else 0, mac.agent.path, mac.seed)
        g = getattr(self, "_graph", None)
        if g is not None and g["key"] == key:
            return g
        T, n, Nn, dev, p = self.episode_limit, self.n_envs, self.n_agents, env.device, N.ptr
        eps_dev = torch.zeros(T, dtype=torch.float32, device=dev)
        rng_dev = torch.zeros(T, dtype=torch.int32, device=dev)
        aios = [N.AgentIO(n_rows=n * Nn, n_steps=1, obs=p(tr["obs"][t]), hidden=p(mac.hidden_states), hidden_zero_init=0,
                          test_mode=int(test_mode), tile_rows=0, path=mac.agent.path, hidden_seq=p(tr["hidden_state"][t]),
                          avail=p(tr["avail_actions"][t]), epsilon=0.0, rng_step=0, seed=mac.seed & 0xFFFFFFFFFFFFFFFF,
                          actions=p(tr["actions_discrete"][t]), power=p(tr["actions_continuous"][t]),
                          epsilon_dev=eps_dev.data_ptr() + 4 * t, rng_step_dev=rng_dev.data_ptr() + 4 * t)
                for t in range(T)]
        if getattr(self, "_structs_for", None) != (mac.hidden_states.data_ptr(), mac.agent.path):
            self._build_step_structs()
        lib, wts = mac.agent.lib(), pk.cstruct()
        graph = torch.cuda.CUDAGraph()
        # (thread-local capture mode: under torch.distributed the NCCL watchdog thread polls its events meanwhile)
        with torch.cuda.graph(graph, capture_error_mode="thread_local"):
            ctx = mac.agent._ctx()                         # the capture stream
            for t in range(T):
                if self.fused_step:
                    lib.call("macjd_rollout_step", ctx, wts, aios[t], env._ctab, self._env_io[t])
                else:
                    lib.call("macjd_agent_forward", ctx, wts, aios[t])
                    lib.call("macjd_env_step", ctx, env._ctab, self._env_io[t])
        self._graph = {"key": key, "graph": graph, "eps_dev": eps_dev, "rng_dev": rng_dev, "keep": (aios, wts),
                       "eps_host": torch.zeros(T, dtype=torch.float32).pin_memory(),
                       "rng_host": torch.zeros(T, dtype=torch.int32).pin_memory()}
        return self._graph

    def _run_graph(self, test_mode):
        mac, T = self.mac, self.episode_limit
        g = self._episode_graph(test_mode)
        eh, rh = g["eps_host"], g["rng_host"]
        for t in range(T):                                 # the same schedule and counters step() would use
            eh[t] = mac.action_selector.anneal(self.t_env + t * self.n_envs, test_mode)
            v = (mac._rng_step + 1 + t) & 0xFFFFFFFF
            rh[t] = v - (1 << 32) if v >= (1 << 31) else v       # uint32 bit pattern in an int32 tensor
        mac._rng_step += T
        self.t_env += T * self.n_envs
        g["eps_dev"].copy_(eh, non_blocking=True)
        g["rng_dev"].copy_(rh, non_blocking=True)
        g["graph"].replay()

    def run(self, test_mode=False, store=True, use_graph=False):
        self.reset()
        if use_graph and self.env.device.type == "cuda" and not getattr(self, "_graph_failed", False):
            try:
                self._run_graph(test_mode)
            except RuntimeError:                           # capture not possible here: keep the per-step launches
                self._graph_failed, use_graph = True, False
                self.reset()
        else:
            use_graph = False
        if not use_graph:
            if self.fused_step and self.whole_episode_launch:
                self.rollout(0, self.episode_limit, test_mode=test_mode)
            else:
                for t in range(self.episode_limit):
                    self.step(t, test_mode=test_mode)
        return self.finish_run(store=store, test_mode=test_mode)

    def finish_run(self, store=True, test_mode=False):
        """Second half of ``run`` (the pipelined training loop issues the timesteps itself): episodes into the
        replay ring, then the reference's ``run_info`` statistics (one read-back)."""
        if store and not test_mode and self.buffer is not None:
            self.buffer.store_rollout(self.traj)
        tr = self.traj
        a = tr["actions_discrete"].view(-1).long()
        counts = torch.bincount(a, minlength=self.args.n_actions)[:self.args.n_actions].float()
        stats = torch.stack([tr["reward"].sum(0).mean(), tr["reward"].mean(), self.r_parts[0].mean(),
                             self.r_parts[1].mean(), self.r_parts[2].mean(), tr["actions_continuous"].mean()]).tolist()
        return {"episode_length": self.episode_limit, "episode_return": stats[0], "avg_step_reward": stats[1],
                "avg_r_d": stats[2], "avg_r_p": stats[3], "avg_r_j": stats[4], "avg_power_overall": stats[5],
                "action_distribution": (counts / counts.sum().clamp(min=1)).cpu().numpy(),
                "n_episodes": self.n_envs}

    def close_env(self):
        self.env.close()
