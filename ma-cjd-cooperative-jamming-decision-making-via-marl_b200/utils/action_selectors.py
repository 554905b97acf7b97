"""Epsilon-greedy action selection (drop-in for utils/action_selectors.py:4-63).

On the hot path the selection is fused into the agent kernel (csrc/agent_act.cuh); this
class owns the epsilon schedule (read by main.py:259,265 through
``mac.action_selector.epsilon``) and offers ``select_action`` on a Q tensor for callers
that use the selector on its own.
"""
import torch


class EpsilonGreedyActionSelector:
    def __init__(self, args):
        self.args = args
        self.epsilon_start = args.epsilon_start
        self.epsilon_finish = args.epsilon_finish
        self.epsilon_anneal_time = args.epsilon_anneal_time
        self.epsilon = self.epsilon_start

    def anneal(self, t_env, test_mode=False):
        """action_selectors.py:30-32 -- linear schedule, frozen in test mode."""
        if not test_mode:
            delta = (self.epsilon_start - self.epsilon_finish) / self.epsilon_anneal_time
            self.epsilon = max(self.epsilon_finish, self.epsilon_start - delta * t_env)
        return self.epsilon

    def select_action(self, agent_qs, avail_actions, t_env, test_mode=False):
        """action_selectors.py:15-63 on tensors [B, N, A] -> long [B, N, 1]."""
        self.anneal(t_env, test_mode)
        masked = agent_qs.masked_fill(avail_actions == 0, -float("inf"))
        greedy = masked.argmax(dim=2)
        if test_mode:
            return greedy.unsqueeze(-1)
        u = torch.rand_like(agent_qs[:, :, 0])
        w = avail_actions.float().clone()
        w[w.sum(dim=-1) == 0] = 1.0 / w.shape[-1]
        rnd = torch.multinomial(w.view(-1, w.shape[-1]), 1).view(agent_qs.shape[0], agent_qs.shape[1])
        pick = (u < self.epsilon).long()
        return (pick * rnd + (1 - pick) * greedy).unsqueeze(-1)
