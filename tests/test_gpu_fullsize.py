"""Full-size parity of the BENCHMARKED configuration (bench.py runs agent_kernel_path = 0: tcgen05 3xTF32 CTA-pair
kernel, split 3 + 4 + 2 unrolls in the learner) -- every check here runs for path 0 as well as for the FP32 SIMT
path 1, with the bound of each quantity written out.

Stated bounds for the tensor-core path (north_star allows a looser one than 1e-5 where TF32 GEMMs are used; the
3xTF32 split keeps ~2^-21 per product, the epilogues use ex2.approx / rcp.approx forms of sigmoid / tanh):
  Q, hidden          rtol 2e-4 of the float64 truth's scale
  learner statistics rtol 5e-4 (loss, grad-norm, Q_tot means)
  gradients          rtol 5e-3 of the largest entry of each tensor
  actions            identical wherever the float64 truth's arg-max margin exceeds 1e-4 x scale; the measured
                     fraction of differing arg-max actions is asserted < 1e-3 and printed.
"""
import types

import numpy as np
import pytest
import torch

from oracle import agent_oracle as AO
from oracle.env_oracle import EnvOracle

pytestmark = pytest.mark.gpu


def _largs(H=128, E=64, A=5, S=24, path=1, lr=5e-6, Nn=2):
    return types.SimpleNamespace(n_agents=Nn, n_actions=A, state_shape=S, obs_shape=S, rnn_hidden_dim=H,
                                 actor_hidden_dim=128, mixing_embed_dim=E, hyper_hidden_dim=128, epsilon_start=1.0,
                                 epsilon_finish=0.05, epsilon_anneal_time=1000, gamma=0.99, lr=lr, grad_norm_clip=1.0,
                                 target_update_interval=200, use_cuda=True, device="cuda", seed=0, agent_kernel_path=path)


def _batch(B, T, Nn, A, S, H, seed):
    rng = np.random.default_rng(seed)
    return {"state": rng.standard_normal((B, T + 1, S)).astype(np.float32),
            "obs": rng.standard_normal((B, T + 1, Nn, S)).astype(np.float32),
            "actions_discrete": rng.integers(0, A, size=(B, T, Nn, 1)).astype(np.int32),
            "actions_continuous": rng.random((B, T, Nn, 1)).astype(np.float32),
            "avail_actions": np.ones((B, T + 1, Nn, A), dtype=np.int64),
            "reward": rng.standard_normal((B, T, 1)).astype(np.float32),
            "terminated": np.zeros((B, T, 1), dtype=bool), "filled": np.ones((B, T, 1), dtype=bool),
            "hidden_state": (rng.standard_normal((B, T + 1, Nn, H)) * 0.5).astype(np.float32), "max_seq_len": T}


def _learner_vs_truth(args, B, T, seed, grads=True):
    """One train step of the CUDA learner against the float64 restatement of core/qmix.py:76-215."""
    from macjd_b200.core.mac import BasicMAC
    from macjd_b200.core.qmix import QMixLearner
    Nn, A, S, H = args.n_agents, args.n_actions, args.state_shape, args.rnn_hidden_dim
    batch = _batch(B, T, Nn, A, S, H, seed)
    torch.manual_seed(42)
    mac = BasicMAC(S, args)
    L = QMixLearner(mac, args)
    agent_sd = {k: v.detach().cpu().clone() for k, v in mac.agent.state_dict().items()}
    mixer_sd = {k: v.detach().cpu().clone() for k, v in L.eval_qmix_net.state_dict().items()}
    stats = L.train(batch, {}, return_debug=True)
    dbg = stats.pop("debug")
    ora = AO.LearnerOracle(agent_sd, mixer_sd, Nn, args.mixing_embed_dim, 0.99, args.lr, 1.0, 200, dtype=torch.float64)
    ref, ref_grads, aux = ora.train({k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in batch.items()})
    return stats, dbg, ref, ref_grads, aux


@pytest.mark.parametrize("path", [1, 0])
def test_learner_c1_full_size_vs_float64_truth(path):
    """B = 32 x T = 100, H = 128, E = 64 (the bench's learner shape) on the benchmarked kernel path."""
    args = _largs(path=path)
    stats, dbg, ref, ref_grads, aux = _learner_vs_truth(args, 32, 100, seed=11)
    tc = path == 0
    for k in ("loss", "grad_norm", "eval_qtot_avg", "target_qtot_avg"):
        np.testing.assert_allclose(stats[k], ref[k], rtol=5e-4 if tc else 2e-4, atol=1e-5, err_msg=k)
    # double-DQN actions (qmix.py:143) against the float64 truth
    mine = dbg["next_actions"].view(99, 32, 2).permute(1, 0, 2).cpu().numpy()
    truth = aux["next_actions"].numpy()
    differ = float((mine != truth).mean())
    print(f"\npath {path}: double-DQN arg-max actions differing from the float64 truth: {differ:.2e} of {truth.size}")
    assert differ < 1e-3
    # un-normalised gradients of every trained tensor: the library keeps sum-gradients, the oracle mean-gradients
    off, denom = 0, float(dbg["sums"][1])
    for nm, n in zip(dbg["names"], dbg["sizes"]):
        g_ref = ref_grads[nm].numpy()
        g = dbg["grad"][off:off + n].cpu().numpy().reshape(g_ref.shape) / denom
        np.testing.assert_allclose(g, g_ref, rtol=0, atol=(5e-3 if tc else 1e-3) * max(1e-12, np.abs(g_ref).max()), err_msg=nm)
        off += n


@pytest.mark.parametrize("path", [1, 0])
def test_learner_c4_stress_shape_vs_float64_truth(path):
    """BASELINE config 4 dims (H = 256, E = 128) at B = 256 x T = 100: statistics against the float64 truth.
    (path 0 = auto: the tensor-core kernel where the dims allow it, else the FP32 kernels.)"""
    args = _largs(H=256, E=128, path=path, lr=1e-4)
    stats, dbg, ref, _, aux = _learner_vs_truth(args, 256, 100, seed=5)
    for k in ("loss", "grad_norm", "eval_qtot_avg", "target_qtot_avg"):
        np.testing.assert_allclose(stats[k], ref[k], rtol=5e-4, atol=1e-5, err_msg=k)
    mine = dbg["next_actions"].view(99, 256, 2).permute(1, 0, 2).cpu().numpy()
    differ = float((mine != aux["next_actions"].numpy()).mean())
    print(f"\npath {path}: H = 256 double-DQN actions differing from the float64 truth: {differ:.2e}")
    assert differ < 1e-3


@pytest.mark.parametrize("path", [1, 0])
def test_c3_rollout_full_size_vs_oracles(path):
    """BASELINE config 3 per-GPU shard: 8 192 envs x 8 jammers x 16 radars x 4 targets (obs 176, 33 actions),
    4 timesteps through BatchedEpisodeRunner with injected env noise and selector draws, against the NumPy
    float64 env oracle and the eager agent oracle step by step."""
    from macjd_b200.simulation.environment import ElectromagneticEnvironment
    from macjd_b200.simulation.scenario import scaled_spec
    from macjd_b200.core.mac import BasicMAC
    from macjd_b200.runners.episode_runner import BatchedEpisodeRunner
    n, J, R, K, T = 8192, 8, 16, 4, 4
    S, A, H = R * 10 + 2 * J, 2 * R + 1, 128
    args = _largs(A=A, S=S, path=path, Nn=J)
    args.episode_limit, args.buffer_size, args.batch_size = T, n, 32
    spec = scaled_spec(n, n_jammers=J, n_radars=R, n_targets=K, seed=21, episode_limit=T)
    env = ElectromagneticEnvironment(args, spec=spec, device="cuda")
    assert env.get_env_info() == {"state_shape": S, "obs_shape": S, "n_actions": A, "n_agents": J, "episode_limit": T}
    args.env_info = env.get_env_info()
    torch.manual_seed(1)
    mac = BasicMAC(S, args)
    mac.cuda()
    runner = BatchedEpisodeRunner(env, mac, None, args)
    ora = EnvOracle(spec)
    sd = {k: v.detach().cpu() for k, v in mac.agent.state_dict().items()}
    sd64 = {k: v.double() for k, v in sd.items()}
    rng = np.random.default_rng(2)
    runner.reset()
    ora.reset()
    np.testing.assert_array_equal(runner.traj["state"][0].cpu().numpy(), ora.get_state())
    h = torch.zeros(n * J, H)
    worst_q, undecidable, flips = 0.0, 0, 0
    for t in range(T):
        noise = rng.random((n, R * K + J)).astype(np.float32)
        u = rng.random((n, J)).astype(np.float32)
        ra = rng.integers(0, A, size=(n, J))
        eps = AO.epsilon_at(runner.t_env, 1.0, 0.05, 1000)
        runner.step(t, noise=torch.from_numpy(noise).cuda(), u_eps=torch.from_numpy(u), rand_actions=torch.from_numpy(ra))
        obs = torch.from_numpy(ora.get_obs())
        # float64 truth of the agent step from the kernel's own previous state (stay on the kernel's trajectory)
        a64, p64, h64, q64, _ = AO.select_actions(sd64, obs.double(), torch.ones(n, J, A, dtype=torch.long), h.double(),
                                                  np.float32(eps), False, torch.from_numpy(u), torch.from_numpy(ra))
        mine_a = runner.traj["actions_discrete"][t].cpu().numpy()
        q = q64.numpy().reshape(n, J, A)
        scale = max(1.0, float(np.abs(q).max()))
        srt = np.sort(q, axis=-1)
        margin = srt[..., -1] - srt[..., -2]
        explore = u < np.float32(eps)
        decidable = (margin > (1e-4 if path == 0 else 1e-5) * scale) | explore
        undecidable += int((~decidable).sum())
        flips += int((mine_a[..., 0] != a64.numpy()[..., 0])[~explore].sum())
        np.testing.assert_array_equal(mine_a[..., 0][decidable], a64.numpy()[..., 0][decidable])
        hk = runner.traj["hidden_state"][t].cpu().reshape(n * J, H)
        tol = 2e-4 if path == 0 else 2e-5
        np.testing.assert_allclose(hk.numpy(), h64.numpy(), rtol=tol, atol=tol)
        # power of the chosen action: compare where the action agrees
        same = mine_a[..., 0] == a64.numpy()[..., 0]
        np.testing.assert_allclose(runner.traj["actions_continuous"][t].cpu().numpy()[..., 0][same], p64.numpy()[..., 0][same],
                                   rtol=tol, atol=tol)
        # env step on the KERNEL's actions: integer outputs bit-exact, float64 reward to 1e-10
        o = ora.step(mine_a.reshape(n, J), runner.traj["actions_continuous"][t].cpu().numpy().reshape(n, J), noise)
        np.testing.assert_array_equal(runner.traj["terminated"][t].cpu().numpy()[:, 0].astype(bool), o["terminated"])
        np.testing.assert_allclose(runner.traj["reward"][t].cpu().numpy()[:, 0], o["reward"], rtol=1e-5, atol=1e-6)
        np.testing.assert_array_equal(runner.traj["state"][t + 1].cpu().numpy(), ora.get_state())
        np.testing.assert_array_equal(runner.traj["avail_actions"][t + 1].cpu().numpy(), np.ones((n, J, A), dtype=np.uint8))
        h = hk.clone()
    print(f"\npath {path}: C3 rollout, {T} x {n * J} agent rows: greedy rows whose float64 margin is below the bound: "
          f"{undecidable}; greedy actions differing from the float64 truth: {flips}")
    assert flips <= undecidable
    assert undecidable < 2e-2 * T * n * J      # (measured on B200: 0.8 % of the rows have a margin below 1e-4 x scale; none flipped)


def _ring_model_slots(size, index, inc):
    """utils/replay_buffer.py:216-250 (reference `_get_storage_idx`) restated on plain integers: the slots the next
    `inc` episodes go to, and the insertion pointer afterwards."""
    if index + inc <= size:
        return np.arange(index, index + inc), index + inc
    if index < size:
        over = inc - (size - index)
        return np.concatenate([np.arange(index, size), np.arange(0, over)]), over
    return np.arange(0, inc), inc


def _episodes_of(gid, T, Nn, S, A, H, dev):
    """Time-major trajectories whose every element is a hash of (global episode number, key, t, inner index): a slot's
    content tells which episode was written there last."""
    n = gid.numel()
    g = gid.to(dev).view(1, n, 1)

    def field(rows, inner, salt, mod):
        t = torch.arange(rows, device=dev, dtype=torch.int64).view(rows, 1, 1)
        j = torch.arange(inner, device=dev, dtype=torch.int64).view(1, 1, inner)
        return (g * 1000003 + t * 7919 + j * 131 + salt) % mod

    return {"state": (field(T + 1, S, 1, 4093).float() - 2046.0) / 64.0,
            "obs": ((field(T + 1, Nn * S, 2, 4093).float() - 2046.0) / 64.0).view(T + 1, n, Nn, S),
            "actions_discrete": field(T, Nn, 3, A).to(torch.int32).view(T, n, Nn, 1),
            "actions_continuous": (field(T, Nn, 4, 1021).float() / 1021.0).view(T, n, Nn, 1),
            "avail_actions": field(T + 1, Nn * A, 5, 2).to(torch.uint8).view(T + 1, n, Nn, A),
            "reward": (field(T, 1, 6, 8191).float() - 4095.0) / 256.0,
            "terminated": field(T, 1, 7, 2).to(torch.uint8),
            "hidden_state": ((field(T + 1, Nn * H, 8, 65521).float() - 32760.0) / 32768.0).view(T + 1, n, Nn, H)}


def test_replay_ring_at_scale_wraps_and_addresses_past_4gb():
    """The C5 ring in the small: 30 000 episodes of the C2 shape (4.3 GB; the hidden_state key alone is 3.1 GB, so slot
    offsets need 64 bits), filled by 9 rollouts of 4 096 episodes (36 864 > 30 000: the ring wraps inside a rollout).
    After every store the pointer and size follow the reference's ring arithmetic (utils/replay_buffer.py:216-250);
    at the end EVERY slot holds, bit for bit and in every key, the episode that was written there last
    (utils/replay_buffer.py:78-156 keeps whole episodes per slot), read back through gather() in both layouts."""
    from macjd_b200.utils.replay_buffer import EpisodeReplayBuffer
    cap, n, T, Nn, S, A, H, dev = 30000, 4096, 100, 2, 24, 5, 128, "cuda"
    args = types.SimpleNamespace(buffer_size=cap, episode_limit=T, n_actions=A, n_agents=Nn, state_shape=S, obs_shape=S,
                                 rnn_hidden_dim=H, use_cuda=True, device=dev)
    buf = EpisodeReplayBuffer(args, device=dev)
    assert buf.buffers["hidden_state"].numel() * buf.buffers["hidden_state"].element_size() > 2 ** 31
    owner = np.full(cap, -1, dtype=np.int64)
    index = size = 0
    for r in range(9):
        gid = torch.arange(r * n, (r + 1) * n, dtype=torch.int64)
        buf.store_rollout(_episodes_of(gid, T, Nn, S, A, H, dev))
        slots, index = _ring_model_slots(cap, index, n)
        size = min(cap, size + n)
        owner[slots] = gid.numpy()
        assert (buf.current_index, buf.current_size) == (index, size), r
    assert (owner >= 0).all() and len(np.unique(owner)) == cap
    rng = np.random.default_rng(0)
    order = rng.permutation(cap)                                   # every slot once, in random order
    for c0 in range(0, cap, 3000):
        idx = order[c0:c0 + 3000]
        want = _episodes_of(torch.from_numpy(owner[idx]), T, Nn, S, A, H, dev)
        tm = (c0 // 3000) % 2 == 1
        got = buf.gather(idx, time_major=tm)
        for k, w in want.items():
            g = got[k] if tm else got[k].transpose(0, 1)
            assert g.shape == w.shape and g.dtype == w.dtype, (k, g.shape, w.shape, g.dtype, w.dtype)
            assert torch.equal(g, w), (k, c0)
        assert bool(got["filled"].all())
        del want, got
