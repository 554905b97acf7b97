"""Multi-agent controller on the fused agent kernel (drop-in for core/mac.py:18-254).

``select_actions`` is one kernel launch (csrc/agent_act.cuh): GRU step + actor + Q-head for
every discrete action + availability mask + epsilon-greedy + gathers.  The random draws of
the selector come from the kernel's Philox stream, or can be injected (``u_eps``,
``rand_actions``) so that the reference and this path can be driven by identical sequences.
"""
from __future__ import annotations

import copy
import os

import numpy as np
import torch

from .networks import RNNAgent
from ..utils.action_selectors import EpsilonGreedyActionSelector


class BasicMAC:
    def __init__(self, input_shape, args, _lib=None):
        self.n_agents = args.n_agents
        self.args = args
        self.input_shape = int(np.prod(input_shape)) if isinstance(input_shape, tuple) else input_shape
        self._lib = _lib
        self._build_agents(self.input_shape, args)
        self.action_selector = EpsilonGreedyActionSelector(args)
        self.hidden_states = None
        self.seed = int(getattr(args, "seed", 0) or 0)
        self._rng_step = 0
        self.last_q_chosen = None
        self._sel_cache = None
        self._host_cache = None

    def __deepcopy__(self, memo):
        # launch-struct caches hold raw addresses of this controller's buffers: never copied
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            setattr(new, k, None if k in ("_sel_cache", "_host_cache") else copy.deepcopy(v, memo))
        return new

    # ------------------------------------------------------------------ acting
    def select_actions(self, obs_batch, avail_actions_batch, t_env, test_mode=False, *, u_eps=None, rand_actions=None):
        """mac.py:59-166.  obs [B, N, obs], avail [B, N, A] -> (long [B, N, 1], float32 [B, N, 1]).
        Updates ``self.hidden_states`` ([B*N, H]) in place."""
        dev = self.device
        B = obs_batch.shape[0]
        M = B * self.n_agents
        if self.hidden_states is None or self.hidden_states.shape[0] != M:
            self.init_hidden(batch_size=B)
        if self.hidden_states.device != dev:
            self.hidden_states = self.hidden_states.to(dev)
        eps = self.action_selector.anneal(t_env, test_mode)
        self._rng_step += 1
        fast = self._select_fast(obs_batch, avail_actions_batch, B, M, eps, test_mode) \
            if (u_eps is None and rand_actions is None) else None
        if fast is not None:
            return fast
        self._rng_step -= 1
        obs = obs_batch.reshape(1, M, self.input_shape)
        avail = avail_actions_batch.reshape(1, M, -1) if avail_actions_batch is not None else None
        self._rng_step += 1
        out = self.agent.run(obs, self.hidden_states, avail=avail, epsilon=eps, test_mode=test_mode,
                             u_eps=u_eps, rand_actions=rand_actions, seed=self.seed, rng_step=self._rng_step,
                             select=True)
        self.last_q_chosen = out["q_chosen"][0]
        actions = out["actions"][0].to(torch.int64).view(B, self.n_agents, 1)
        power = out["power"][0].view(B, self.n_agents, 1)
        return actions, power

    def _select_fast(self, obs, avail, B, M, eps, test_mode):
        """Steady-state path: when the caller keeps handing in the same device buffers (a rollout
        loop does), the launch struct and the output tensors are cached and a call costs one ctypes
        invocation.  Falls back (returns None) for anything unusual."""
        from .. import _native as N
        if not (torch.is_tensor(obs) and torch.is_tensor(avail) and obs.device == self.device and avail.device == self.device
                and obs.dtype == torch.float32 and avail.dtype == torch.uint8 and obs.is_contiguous() and avail.is_contiguous()):
            return None
        key = (M, obs.data_ptr(), avail.data_ptr(), self.hidden_states.data_ptr(), self.agent.path)
        c = self._sel_cache
        if c is None or c["key"] != key:
            dev = self.device
            out = {"actions": torch.empty(M, dtype=torch.int32, device=dev), "power": torch.empty(M, dtype=torch.float32, device=dev),
                   "q_chosen": torch.empty(M, dtype=torch.float32, device=dev)}
            io = N.AgentIO(n_rows=M, n_steps=1, obs=obs.data_ptr(), hidden=self.hidden_states.data_ptr(), hidden_zero_init=0,
                           test_mode=0, tile_rows=0, path=self.agent.path, avail=avail.data_ptr(), epsilon=0.0, rng_step=0,
                           seed=self.seed & 0xFFFFFFFFFFFFFFFF, actions=out["actions"].data_ptr(),
                           power=out["power"].data_ptr(), q_chosen=out["q_chosen"].data_ptr())
            c = self._sel_cache = {"key": key, "io": io, "out": out, "keep": (obs, avail)}
        io = c["io"]
        io.epsilon, io.rng_step, io.test_mode = float(eps), self._rng_step & 0xFFFFFFFF, int(test_mode)
        self.agent.lib().call("macjd_agent_forward", self.agent._ctx(), self.agent.packed().cstruct(), io)
        out = c["out"]
        self.last_q_chosen = out["q_chosen"]
        return out["actions"].to(torch.int64).view(B, self.n_agents, 1), out["power"].clone().view(B, self.n_agents, 1)

    def _host_launch(self, obs, avail, actions_out=None, power_out=None):
        """Launch struct + host-pointer struct for the host-buffer calls, cached on the identity of the
        caller's buffers (a rollout loop hands in the same objects every step)."""
        from .. import _native as N
        # steady state: the caller hands in the same buffer objects every step -> identity checks only
        c = self._host_cache
        if not (c is not None and c["keep"][0] is obs and c["keep"][1] is avail and (actions_out is None or c["keep"][2] is actions_out)
                and (power_out is None or c["keep"][3] is power_out) and c["hid"] is self.hidden_states and c["path"] == self.agent.path):
            B = obs.shape[0]
            M = B * self.n_agents
            dev = self.device
            # obs [B, obs]: ONE row per env -- the global state, which every agent of the env observes in this
            # environment (environment.py:512-522 returns get_state() once per jammer); the launch then runs with
            # obs_group = n_agents and half as many bytes cross PCIe at two jammers.  obs [B, N, obs]: the reference's layout.
            shared = obs.dim() == 2 if torch.is_tensor(obs) else np.ndim(obs) == 2
            if self.hidden_states is None or self.hidden_states.shape[0] != M:
                self.init_hidden(batch_size=B)
            if self.hidden_states.device != dev:
                self.hidden_states = self.hidden_states.to(dev)
            A = self.args.n_actions
            pin = dev.type == "cuda"
            if actions_out is None:
                actions_out = torch.zeros(B, self.n_agents, dtype=torch.int32, pin_memory=pin)
            if power_out is None:
                power_out = torch.zeros(B, self.n_agents, dtype=torch.float32, pin_memory=pin)
            st = {"obs": torch.empty(B if shared else M, self.input_shape, dtype=torch.float32, device=dev),
                  "avail": torch.ones(M, A, dtype=torch.uint8, device=dev),
                  "actions": torch.empty(M, dtype=torch.int32, device=dev), "power": torch.empty(M, dtype=torch.float32, device=dev),
                  "q_chosen": torch.empty(M, dtype=torch.float32, device=dev)}
            io = N.AgentIO(n_rows=M, n_steps=1, obs=st["obs"].data_ptr(), hidden=self.hidden_states.data_ptr(), hidden_zero_init=0,
                           test_mode=0, tile_rows=0, path=self.agent.path, avail=st["avail"].data_ptr(), epsilon=0.0, rng_step=0,
                           seed=self.seed & 0xFFFFFFFFFFFFFFFF, actions=st["actions"].data_ptr(),
                           power=st["power"].data_ptr(), q_chosen=st["q_chosen"].data_ptr(),
                           obs_group=self.n_agents if shared else 0)
            pinned = dev.type == "cuda" and all(torch.is_tensor(v) and v.is_pinned() for v in (obs, avail, actions_out, power_out))
            hs = N.ActHost(obs=N.ptr(obs), avail=N.ptr(avail), actions=N.ptr(actions_out), power=N.ptr(power_out), q_chosen=None,
                           flags=N.HOST_PINNED if pinned else 0)
            c = self._host_cache = {"io": io, "host": hs, "stage": st, "keep": (obs, avail, actions_out, power_out),
                                    "hid": self.hidden_states, "path": self.agent.path}
        return c

    def select_actions_host(self, obs, avail, t_env, test_mode=False, *, actions_out=None, power_out=None):
        """mac.py:59-166 for callers that live on the host, as the reference's runner does: ``obs``
        float32 [B, N, obs] and ``avail`` uint8 [B, N, A] are contiguous CPU tensors / numpy arrays
        (page-locked memory makes the copies asynchronous); returns the chosen discrete actions int32
        [B, N] and power float32 [B, N] in host memory (``actions_out`` / ``power_out`` when given).
        One C call: copies in, fused step, copies out, stream drained on return
        (include/macjd.h: macjd_agent_act_host).  The recurrent state stays on the device."""
        c = self._host_launch(obs, avail, actions_out, power_out)
        eps = self.action_selector.anneal(t_env, test_mode)
        self._rng_step += 1
        io = c["io"]
        io.epsilon, io.rng_step, io.test_mode = float(eps), self._rng_step & 0xFFFFFFFF, int(test_mode)
        self.agent.lib().call("macjd_agent_act_host", self.agent._ctx(), self.agent.packed().cstruct(), io, c["host"])
        self.last_q_chosen = c["stage"]["q_chosen"]
        return c["keep"][2], c["keep"][3]

    def host_step_args(self, obs, avail, t_env, test_mode=False, *, actions_out=None, power_out=None):
        """The (weights, io, host) structs of one ``select_actions_host`` call without making it: the
        runner's fused act + env step (BatchedEpisodeRunner.step_host) passes them to
        macjd_rollout_step_host.  Advances epsilon / the Philox step exactly as the call would."""
        c = self._host_launch(obs, avail, actions_out, power_out)
        eps = self.action_selector.anneal(t_env, test_mode)
        self._rng_step += 1
        io = c["io"]
        io.epsilon, io.rng_step, io.test_mode = float(eps), self._rng_step & 0xFFFFFFFF, int(test_mode)
        self.last_q_chosen = c["stage"]["q_chosen"]
        return self.agent.packed().cstruct(), io, c["host"], c["keep"][2], c["keep"][3]

    def forward(self, agent_inputs_reshaped, hidden_states):
        """mac.py:168-187 -> (h_out [M, H], continuous_params_all [M, A])."""
        h = hidden_states.detach().to(dtype=torch.float32).contiguous().clone()
        out = self.agent.run(agent_inputs_reshaped, h, want_params=True)
        return out["hidden"], out["params_all"][0]

    def init_hidden(self, batch_size):
        """mac.py:189-198"""
        self.hidden_states = torch.zeros(batch_size * self.n_agents, self.args.rnn_hidden_dim,
                                         dtype=torch.float32, device=self.device)

    # ------------------------------------------------------------------ utilities
    @property
    def device(self):
        return self.agent.fc1.weight.device

    def parameters(self):
        return self.agent.parameters()

    def load_state(self, other_mac_state_dict):
        self.agent.load_state_dict(other_mac_state_dict)

    def state_dict(self):
        return self.agent.state_dict()

    def cuda(self):
        self.agent.cuda()
        if self.hidden_states is not None:
            self.hidden_states = self.hidden_states.cuda()

    def save_models(self, path):
        os.makedirs(path, exist_ok=True)
        torch.save(self.agent.state_dict(), f"{path}/agent.pth")

    def load_models(self, path):
        self.agent.load_state_dict(torch.load(f"{path}/agent.pth", map_location=self.device))

    def _build_agents(self, input_shape, args):
        self.agent = RNNAgent(input_shape, args, _lib=self._lib)
