"""Eager-PyTorch oracle of the MP-DQN agent, the controller's action selection, the QMix
mixer and the QMix learner step.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Functional restatement working directly on ``state_dict`` tensors (the reference's key
names), in float32 by default and in float64 on request (tie analysis).

Follows (reference paths relative to /root/reference):
  * core/networks.py:88-114   RNNAgent.forward: h' = GRUCell(relu(fc1 x), h)
  * core/networks.py:116-129  actor_forward: P = sigmoid(MLP3(x))
  * core/networks.py:131-180  get_q_value_for_action: W2 relu(W1 [h; onehot(a); P_a] + b1) + b2
  * core/mac.py:59-166        select_actions (A sequential Q passes, mask, selector, gather)
  * utils/action_selectors.py:15-63  epsilon schedule, mask, Bernoulli(eps), random
                                     available action, first-max argmax, test-mode override
  * core/networks.py:250-316  QMixer.forward (LayerNorm, 4 hypernets, clamp, bmm, ELU)
  * core/qmix.py:76-215       QMixLearner.train (two unrolls, double-DQN target, q_taken on
                              STORED hidden states, masked MSE, clip, Adam, hard target sync)
Third-party arithmetic restated from its published definition: torch.nn.GRUCell (gate
order r, z, n), torch.nn.LayerNorm (eps 1e-5, biased variance), F.elu (alpha 1),
clip_grad_norm_ (coef = max_norm / (norm + 1e-6), clamped to 1), Adam (betas 0.9/0.999,
eps 1e-8, bias-corrected).

Parity status: PINNED against the unmodified reference modules run in the build container
(tests/golden/agent_*.npz, mixer_*.npz, learner_*.npz from tests/golden/make_golden.py).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _lin(sd, name, x):
    return x @ sd[name + ".weight"].t() + sd[name + ".bias"]


def cast_sd(sd, dtype):
    return {k: v.detach().to(dtype) for k, v in sd.items()}


# ------------------------------------------------------------------ agent (a6)
def gru_cell(sd, x, h):
    """torch.nn.GRUCell restated: rows of weight_ih / weight_hh are the r, z, n blocks."""
    gi = x @ sd["rnn.weight_ih"].t() + sd["rnn.bias_ih"]
    gh = h @ sd["rnn.weight_hh"].t() + sd["rnn.bias_hh"]
    H = h.shape[1]
    r = torch.sigmoid(gi[:, :H] + gh[:, :H])
    z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
    n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
    return (1.0 - z) * n + z * h


def agent_hidden(sd, obs, h):
    """core/networks.py:88-114"""
    return gru_cell(sd, torch.relu(_lin(sd, "fc1", obs)), h)


def actor_params(sd, obs):
    """core/networks.py:116-129"""
    a = torch.relu(_lin(sd, "actor.0", obs))
    a = torch.relu(_lin(sd, "actor.2", a))
    return torch.sigmoid(_lin(sd, "actor.4", a))


def q_for_action(sd, h, a_idx, p):
    """core/networks.py:131-180.  h [M,H], a_idx long [M], p [M] -> q [M]."""
    n_actions = sd["actor.4.weight"].shape[0]
    if torch.any(a_idx < 0) or torch.any(a_idx >= n_actions):
        raise IndexError("Action index out of bounds")
    onehot = F.one_hot(a_idx.long(), n_actions).to(h.dtype)
    inp = torch.cat([h, onehot, p.reshape(-1, 1)], dim=1)
    hid = torch.relu(_lin(sd, "fc2_q_head.0", inp))
    return _lin(sd, "fc2_q_head.2", hid).squeeze(1)


def q_all_actions(sd, h, params):
    """core/mac.py:112-135 / core/qmix.py:260-274 -- one Q-head pass per discrete action."""
    M, A = params.shape
    cols = []
    for a in range(A):
        cols.append(q_for_action(sd, h, torch.full((M,), a, dtype=torch.long), params[:, a]))
    return torch.stack(cols, dim=1)


# ------------------------------------------------------------------ selector (a7) / MAC (a5)
def epsilon_at(t_env, start, finish, anneal):
    """utils/action_selectors.py:30-32"""
    delta = (start - finish) / anneal
    return max(finish, start - delta * t_env)


def select_actions(sd, obs, avail, h, eps, test_mode, u, rand_actions):
    """core/mac.py:59-166 with the selector's random draws injected.
    obs [B,N,O]; avail [B,N,A] (0/1); h [B*N,H]; u [B,N] uniforms; rand_actions long [B,N].
    Returns (actions long [B,N,1], power [B,N,1], h' [B*N,H], q_masked [B,N,A], params [B*N,A])."""
    B, Nn, O = obs.shape
    x = obs.reshape(B * Nn, O)
    h2 = agent_hidden(sd, x, h)
    params = actor_params(sd, x)
    q = q_all_actions(sd, h2, params).view(B, Nn, -1).clone()
    q[avail == 0] = -float("inf")
    greedy = q.argmax(dim=2)
    if test_mode:
        chosen = greedy
    else:
        pick = (u < eps).long()
        chosen = pick * rand_actions + (1 - pick) * greedy
    power = torch.gather(params, 1, chosen.view(-1, 1)).view(B, Nn, 1)
    return chosen.unsqueeze(-1), power, h2, q, params


# ------------------------------------------------------------------ mixer (a8)
def layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def mixer_forward(msd, agent_qs, states, n_agents, embed_dim):
    """core/networks.py:250-316.  agent_qs [R,N]; states [R,S] -> q_tot [R]."""
    s = layer_norm(states, msd["state_norm.weight"], msd["state_norm.bias"])
    w1 = torch.clamp(_lin(msd, "hyper_w_1.2", torch.relu(_lin(msd, "hyper_w_1.0", s))), 0.0, 5.0)
    b1 = torch.clamp(_lin(msd, "hyper_b_1", s), -5.0, 5.0)
    wf = torch.clamp(_lin(msd, "hyper_w_final.2", torch.relu(_lin(msd, "hyper_w_final.0", s))), 0.0, 5.0)
    v = torch.clamp(_lin(msd, "V.2", torch.relu(_lin(msd, "V.0", s))), -5.0, 5.0)
    w1 = w1.view(-1, n_agents, embed_dim)
    hidden = F.elu(torch.bmm(agent_qs.view(-1, 1, n_agents), w1).squeeze(1) + b1)
    return (hidden * wf).sum(dim=1) + v.squeeze(1)


# ------------------------------------------------------------------ learner (a9)
def unroll_q(sd, obs, n_agents):
    """core/qmix.py:217-280.  obs [B,T,N,O] -> Q for all actions [B,T,N,A] from a zero
    initial hidden state (the eval / target unroll)."""
    B, T, Nn, O = obs.shape
    H = sd["rnn.weight_hh"].shape[1]
    h = torch.zeros(B * Nn, H, dtype=obs.dtype)
    out = []
    for t in range(T):
        x = obs[:, t].reshape(B * Nn, O)
        h = agent_hidden(sd, x, h)
        out.append(q_all_actions(sd, h, actor_params(sd, x)).view(B, Nn, -1))
    return torch.stack(out, dim=1)


def td_loss(agent_sd, mixer_sd, tgt_agent_sd, tgt_mixer_sd, batch, gamma, n_agents, embed_dim):
    """core/qmix.py:98-194 up to the loss.  ``batch``: dict of tensors in the reference's
    sampled-batch layout, already trimmed (state/obs/hidden_state have T+1 slots)."""
    T = int(batch["max_seq_len"])
    states, obs = batch["state"][:, :T], batch["obs"][:, :T]
    a_d, a_c = batch["actions_discrete"][:, :T].long(), batch["actions_continuous"][:, :T]
    rewards, term = batch["reward"][:, :T], batch["terminated"][:, :T].to(states.dtype)
    mask = batch["filled"][:, :T].to(states.dtype).squeeze(-1)
    hidden = batch["hidden_state"][:, :T + 1]
    B = states.shape[0]
    with torch.no_grad():
        tq = unroll_q(tgt_agent_sd, obs, n_agents)
        eq = unroll_q(agent_sd, obs, n_agents)
        nxt = eq[:, 1:].argmax(dim=3, keepdim=True)            # no avail mask (qmix.py:141-143)
        tq_taken = torch.gather(tq[:, 1:], 3, nxt).squeeze(3)  # [B,T-1,N]
        tq_tot = mixer_forward(tgt_mixer_sd, tq_taken.reshape(-1, n_agents),
                               states[:, 1:].reshape(B * (T - 1), -1), n_agents, embed_dim).view(B, T - 1, 1)
        targets = rewards[:, :-1] + gamma * (1 - term[:, :-1]) * tq_tot
    Hd = hidden.shape[-1]
    q_taken = q_for_action(agent_sd, hidden[:, :T - 1].reshape(-1, Hd),
                           a_d[:, :T - 1].reshape(-1), a_c[:, :T - 1].reshape(-1)).view(B, T - 1, n_agents)
    q_tot = mixer_forward(mixer_sd, q_taken.reshape(-1, n_agents),
                          states[:, :-1].reshape(B * (T - 1), -1), n_agents, embed_dim).view(B, T - 1, 1)
    td = q_tot - targets
    m = mask[:, :-1]
    loss = ((td * m.unsqueeze(-1)) ** 2).sum() / m.sum()
    return loss, {"q_tot": q_tot, "targets": targets, "q_taken": q_taken, "tq_taken": tq_taken,
                  "next_actions": nxt.squeeze(3)}


TRAINED_AGENT_KEYS = ("fc2_q_head.0.weight", "fc2_q_head.0.bias", "fc2_q_head.2.weight", "fc2_q_head.2.bias")


class LearnerOracle:
    """QMixLearner restated (core/qmix.py:25-215): only the Q-head and the eval mixer
    receive gradients, so only they have Adam state."""

    def __init__(self, agent_sd, mixer_sd, n_agents, embed_dim, gamma, lr, grad_norm_clip,
                 target_update_interval, dtype=torch.float32):
        self.dtype = dtype
        self.agent = cast_sd(agent_sd, dtype)
        self.mixer = cast_sd(mixer_sd, dtype)
        self.tgt_agent = {k: v.clone() for k, v in self.agent.items()}
        self.tgt_mixer = {k: v.clone() for k, v in self.mixer.items()}
        self.n_agents, self.embed_dim = n_agents, embed_dim
        self.gamma, self.lr, self.clip, self.interval = gamma, lr, grad_norm_clip, target_update_interval
        self.train_step = 0
        self.last_target_update_step = 0
        self.adam = {}     # name -> (step, m, v)

    def trainable(self):
        out = [("agent." + k, self.agent, k) for k in TRAINED_AGENT_KEYS]
        out += [("mixer." + k, self.mixer, k) for k in self.mixer]
        return out

    def train(self, batch):
        self.train_step += 1
        batch = {k: (v.to(self.dtype) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in batch.items()}
        leaves = []
        for _, sd, k in self.trainable():
            sd[k] = sd[k].detach().requires_grad_(True)
            leaves.append(sd[k])
        loss, aux = td_loss(self.agent, self.mixer, self.tgt_agent, self.tgt_mixer, batch,
                            self.gamma, self.n_agents, self.embed_dim)
        grads = torch.autograd.grad(loss, leaves)
        total = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads))
        coef = min(1.0, self.clip / (total + 1e-6))
        grad_dict = {}
        with torch.no_grad():
            for (name, sd, k), g in zip(self.trainable(), grads):
                grad_dict[name] = g.clone()
                g = g * coef
                step, m, v = self.adam.get(name, (0, torch.zeros_like(g), torch.zeros_like(g)))
                step += 1
                m = 0.9 * m + 0.1 * g
                v = 0.999 * v + 0.001 * g * g
                denom = v.sqrt() / math.sqrt(1 - 0.999 ** step) + 1e-8
                sd[k] = (sd[k] - (self.lr / (1 - 0.9 ** step)) * m / denom).detach()
                self.adam[name] = (step, m, v)
        if self.train_step - self.last_target_update_step >= self.interval:
            self.tgt_agent = {k: v.detach().clone() for k, v in self.agent.items()}
            self.tgt_mixer = {k: v.detach().clone() for k, v in self.mixer.items()}
            self.last_target_update_step = self.train_step
        stats = {"loss": float(loss.detach()), "grad_norm": total,
                 "eval_qtot_avg": float(aux["q_tot"].mean()), "target_qtot_avg": float(aux["targets"].mean())}
        return stats, grad_dict, aux
