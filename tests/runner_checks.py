"""Runner checks shared by the host-emulation (CPU) and GPU tests."""
import os
import tempfile
import types

import numpy as np
import pytest
import torch
import yaml

from oracle.env_oracle import EnvOracle
from oracle import agent_oracle as AO


def rl_args(device, **kw):
    base = dict(n_agents=2, n_actions=5, state_shape=24, obs_shape=24, rnn_hidden_dim=64, actor_hidden_dim=64,
                mixing_embed_dim=32, hyper_hidden_dim=64, epsilon_start=1.0, epsilon_finish=0.05,
                epsilon_anneal_time=50, gamma=0.99, lr=1e-3, grad_norm_clip=1.0, target_update_interval=2,
                use_cuda=torch.device(device).type == "cuda", device=device, batch_size=4, buffer_size=16,
                episode_limit=6, seed=3)
    base.update(kw)
    a = types.SimpleNamespace(**base)
    a.env_info = {"state_shape": a.state_shape, "obs_shape": a.obs_shape, "n_actions": a.n_actions,
                  "n_agents": a.n_agents, "episode_limit": a.episode_limit}
    return a


def check_batched_rollout_against_oracles(device, lib, n_envs=9):
    """A full batched rollout with injected env noise and selector draws, replayed step by step
    through the NumPy env oracle and the eager agent oracle: actions, rewards, stored hidden
    states and the replay contents must agree."""
    from macjd_b200.simulation.environment import ElectromagneticEnvironment
    from macjd_b200.simulation.scenario import hetero_spec
    from macjd_b200.core.mac import BasicMAC
    from macjd_b200.utils.replay_buffer import EpisodeReplayBuffer
    from macjd_b200.runners.episode_runner import BatchedEpisodeRunner
    args = rl_args(device, buffer_size=max(16, 2 * n_envs))
    spec = hetero_spec(n_envs, seed=9, active=True, episode_limit=args.episode_limit)
    env = ElectromagneticEnvironment(args, spec=spec, device=device, _lib=lib)
    torch.manual_seed(1)
    mac = BasicMAC(24, args, _lib=lib)
    if args.use_cuda:
        mac.cuda()
    buf = EpisodeReplayBuffer(args, device=device, _lib=lib)
    runner = BatchedEpisodeRunner(env, mac, buf, args)
    ora = EnvOracle(spec)
    sd = {k: v.detach().cpu() for k, v in mac.agent.state_dict().items()}
    rng = np.random.default_rng(2)
    T, Nn = args.episode_limit, 2
    runner.reset()
    ora.reset()
    h = torch.zeros(n_envs * Nn, args.rnn_hidden_dim)
    dev = env.device
    for t in range(T):
        noise = rng.random((n_envs, 4)).astype(np.float32)
        u = rng.random((n_envs, Nn)).astype(np.float32)
        ra = rng.integers(0, 5, size=(n_envs, Nn))
        eps = AO.epsilon_at(runner.t_env, 1.0, 0.05, 50)
        runner.step(t, noise=torch.from_numpy(noise).to(dev), u_eps=torch.from_numpy(u), rand_actions=torch.from_numpy(ra))
        obs = torch.from_numpy(ora.get_obs())
        a, p, h, q, _ = AO.select_actions(sd, obs, torch.ones(n_envs, Nn, 5, dtype=torch.long), h, np.float32(eps), False,
                                          torch.from_numpy(u), torch.from_numpy(ra))
        mine_a = runner.traj["actions_discrete"][t].cpu().numpy()
        srt = np.sort(q.numpy(), axis=-1)
        decidable = ((srt[..., -1] - srt[..., -2]) > 1e-5)[..., None] | (u < np.float32(eps))[..., None]
        assert decidable.all(), "test inputs should have decidable argmaxes"
        np.testing.assert_array_equal(mine_a, a.numpy())
        np.testing.assert_allclose(runner.traj["actions_continuous"][t].cpu().numpy(), p.numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(runner.traj["hidden_state"][t].cpu().numpy().reshape(n_envs * Nn, -1), h.numpy(), rtol=1e-4, atol=1e-5)
        o = ora.step(mine_a.reshape(n_envs, Nn), runner.traj["actions_continuous"][t].cpu().numpy().reshape(n_envs, Nn), noise)
        np.testing.assert_allclose(runner.traj["reward"][t].cpu().numpy()[:, 0], o["reward"], rtol=1e-5, atol=1e-6)
        np.testing.assert_array_equal(runner.traj["terminated"][t].cpu().numpy()[:, 0].astype(bool), o["terminated"])
        h = runner.traj["hidden_state"][t].cpu().reshape(n_envs * Nn, -1).clone()    # stay on the kernel's trajectory
    buf.store_rollout(runner.traj)
    got = buf.gather(np.arange(n_envs))
    for k in ("state", "obs", "reward", "hidden_state", "actions_discrete"):
        assert torch.equal(got[k].transpose(0, 1), runner.traj[k]), k
    assert got["max_seq_len"] == T
    return runner, buf, args


def check_training_loop_smoke(device, lib):
    """act -> store -> sample -> train -> target update on the batched path; finite stats,
    the Q-head moves, everything else of the agent stays frozen."""
    from macjd_b200.core.qmix import QMixLearner
    runner, buf, args = check_batched_rollout_against_oracles(device, lib, n_envs=6)
    learner = QMixLearner(runner.mac, args, _lib=lib)
    before = {k: v.detach().cpu().clone() for k, v in runner.mac.agent.state_dict().items()}
    info = runner.run()
    assert info["episode_length"] == args.episode_limit and np.isfinite(info["episode_return"])
    assert abs(info["action_distribution"].sum() - 1) < 1e-6
    np.random.seed(0)
    for it in range(3):
        stats = learner.train(buf.sample(args.batch_size, time_major=(it % 2 == 0)), {})
        assert all(np.isfinite(v) for v in stats.values()), stats
    after = runner.mac.agent.state_dict()
    for k in before:
        changed = not torch.equal(before[k], after[k].cpu())
        assert changed == k.startswith("fc2_q_head"), k
    assert learner.last_target_update_step == 2


def check_reference_protocol_runner(device, lib):
    """EpisodeRunner (one env, host staging) against the shim environment: stored episode has
    the reference's shapes / dtypes and the zero last slot."""
    from macjd_b200.simulation.environment import ElectromagneticEnvironment
    from macjd_b200.simulation.scenario import default_config_dict
    from macjd_b200.core.mac import BasicMAC
    from macjd_b200.utils.replay_buffer import EpisodeReplayBuffer
    from macjd_b200.runners.episode_runner import EpisodeRunner
    args = rl_args(device, episode_limit=5)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "simulation_config.yaml")
        with open(path, "w") as f:
            yaml.safe_dump(default_config_dict(), f)
        env = ElectromagneticEnvironment(args, sim_config_path=path, device=device, _lib=lib)
    mac = BasicMAC(24, args, _lib=lib)
    if args.use_cuda:
        mac.cuda()
    buf = EpisodeReplayBuffer(args, device=device, _lib=lib)
    runner = EpisodeRunner(env, mac, buf, args)
    info = runner.run()
    assert info["episode_length"] == 5 and runner.t_env == 5 and len(buf) == 1
    assert set(info) >= {"episode_length", "episode_return", "avg_step_reward", "avg_r_d", "avg_r_p", "avg_r_j",
                         "avg_power_overall", "action_distribution"}
    ep = buf.gather(np.array([0]))
    assert ep["max_seq_len"] == 5
    assert float(ep["state"][0, 5].abs().sum()) == 0.0 and float(ep["hidden_state"][0, 5].abs().sum()) == 0.0
    assert float(ep["state"][0, 4].abs().sum()) > 0 and bool(ep["filled"].all())
    assert not bool(ep["terminated"][0, :4].any()) and bool(ep["terminated"][0, 4].all())
    runner.close_env()


def check_host_buffer_api(device, lib, n_envs=7, steps=5, pinned=True):
    """The host-buffer entry points (macjd_agent_act_host / macjd_env_step_host: numpy in, numpy out,
    copies inside the call) against the device-resident API on twin envs / controllers driven greedily:
    identical actions, power, rewards, terminated flags and observations at every step."""
    import copy
    from macjd_b200.simulation.environment import ElectromagneticEnvironment
    from macjd_b200.simulation.scenario import hetero_spec
    from macjd_b200.core.mac import BasicMAC
    args = rl_args(device)
    spec = hetero_spec(n_envs, seed=4, active=True, episode_limit=steps - 1)
    envs = [ElectromagneticEnvironment(args, spec=spec, device=device, seed=11, _lib=lib) for _ in range(2)]
    torch.manual_seed(5)
    mac_a = BasicMAC(24, args, _lib=lib)
    if args.use_cuda:
        mac_a.cuda()
    mac_a.select_actions(envs[0].get_obs(), envs[0].get_avail_actions(), 0, test_mode=True)   # fills the launch caches
    mac_b = copy.deepcopy(mac_a)                  # a controller with live caches must stay copyable
    for m in (mac_a, mac_b):
        m.init_hidden(n_envs)
        m._rng_step = 0
    env_d, env_h = envs
    hb = env_h.host_buffers(pinned=pinned)   # pinned + small: kernels work on host memory directly; else copy engines
    obs_h = env_h.get_obs().cpu().contiguous()
    avail_h = env_h.get_avail_actions().cpu().contiguous()
    for t in range(steps):
        a_d, p_d = mac_a.select_actions(env_d.get_obs(), env_d.get_avail_actions(), t, test_mode=(t % 2 == 0))
        a_h, p_h = mac_b.select_actions_host(obs_h, avail_h, t, test_mode=(t % 2 == 0),
                                             actions_out=hb["act_d"], power_out=hb["act_p"])
        assert a_h is hb["act_d"] and p_h is hb["act_p"]
        np.testing.assert_array_equal(a_h.numpy(), a_d.cpu().numpy()[..., 0])
        np.testing.assert_array_equal(p_h.numpy(), p_d.cpu().numpy()[..., 0])
        obs_d, rew_d, term_d, _ = env_d.step((a_d[..., 0], p_d[..., 0]))
        env_h.step_host(hb)
        np.testing.assert_array_equal(hb["reward"].numpy(), rew_d.cpu().numpy())
        np.testing.assert_array_equal(hb["terminated"].numpy().astype(bool), term_d.cpu().numpy())
        np.testing.assert_array_equal(hb["obs"].numpy(), obs_d.cpu().numpy())
        obs_h = hb["obs"]
    with pytest.raises(ValueError):
        env_h.step_host({"act_d": hb["act_d"]})


def check_fused_host_step(device, lib, n_envs=9, steps=6, pinned=True):
    """BatchedEpisodeRunner.step_host (macjd_rollout_step_host: one call, one drain) against
    select_actions_host + step_host on twin envs / controllers, exploration draws included."""
    import copy
    import types
    from macjd_b200.simulation.environment import ElectromagneticEnvironment
    from macjd_b200.simulation.scenario import hetero_spec
    from macjd_b200.core.mac import BasicMAC
    from macjd_b200.runners.episode_runner import BatchedEpisodeRunner
    args = rl_args(device, epsilon_anneal_time=20)
    spec = hetero_spec(n_envs, seed=6, active=True, episode_limit=steps - 2)
    envs = [ElectromagneticEnvironment(args, spec=spec, device=device, seed=13, _lib=lib) for _ in range(3)]
    torch.manual_seed(9)
    mac_a = BasicMAC(24, args, _lib=lib)
    if args.use_cuda:
        mac_a.cuda()
    mac_b, mac_c = copy.deepcopy(mac_a), copy.deepcopy(mac_a)
    for m in (mac_a, mac_b, mac_c):
        m.init_hidden(n_envs)
    runner = types.SimpleNamespace(mac=mac_b, env=envs[1], t_env=0)
    # third twin: the host keeps ONE observation row per env (the state) instead of the per-jammer copies
    runner_s = types.SimpleNamespace(mac=mac_c, env=envs[2], t_env=0)
    hbs = [e.host_buffers(pinned=pinned) for e in envs]
    for hb, e in zip(hbs, envs):
        hb["state"] = torch.zeros(n_envs, e.state_dim)
        if pinned and torch.device(device).type == "cuda":
            hb["state"] = hb["state"].pin_memory()
        e.reset()
    obs = [e.get_obs().cpu().contiguous() for e in envs]
    avail = [e.get_avail_actions().cpu().contiguous() for e in envs]
    hb_s = {k: v for k, v in hbs[2].items() if k != "obs"}
    hb_s["state"].copy_(envs[2].get_state().cpu())
    for t in range(steps):
        mac_a.select_actions_host(obs[0], avail[0], t * n_envs, actions_out=hbs[0]["act_d"], power_out=hbs[0]["act_p"])
        envs[0].step_host(hbs[0])
        BatchedEpisodeRunner.step_host(runner, obs[1], avail[1], hbs[1])
        BatchedEpisodeRunner.step_host(runner_s, hb_s["state"], avail[2], hb_s)
        for k in hbs[0]:
            np.testing.assert_array_equal(hbs[1][k].numpy(), hbs[0][k].numpy(), err_msg=f"{k} t={t}")
            if k != "obs":
                np.testing.assert_array_equal(hb_s[k].numpy(), hbs[0][k].numpy(), err_msg=f"state-only {k} t={t}")
        assert torch.equal(mac_a.hidden_states, mac_b.hidden_states) and torch.equal(mac_a.hidden_states, mac_c.hidden_states)
        obs = [hbs[0]["obs"], hbs[1]["obs"]]
    assert runner.t_env == steps * n_envs       # the reference's unit: single-environment steps


def check_episode_graph_equals_stepwise(device, n_envs=64):
    """BatchedEpisodeRunner.run replays a whole episode as one CUDA graph (epsilon and the Philox counter
    come from device memory); twin runners, one stepping launch by launch, must produce identical
    trajectories over consecutive episodes -- exploration draws included."""
    import copy
    from macjd_b200.simulation.environment import ElectromagneticEnvironment
    from macjd_b200.simulation.scenario import hetero_spec
    from macjd_b200.core.mac import BasicMAC
    from macjd_b200.utils.replay_buffer import EpisodeReplayBuffer
    from macjd_b200.runners.episode_runner import BatchedEpisodeRunner
    args = rl_args(device, buffer_size=4 * n_envs, rnn_hidden_dim=128, actor_hidden_dim=128, agent_kernel_path=0,
                   epsilon_anneal_time=40)
    spec = hetero_spec(n_envs, seed=2, active=True, episode_limit=args.episode_limit)
    torch.manual_seed(8)
    mac0 = BasicMAC(24, args)
    mac0.cuda()
    runners = []
    for _ in range(2):
        env = ElectromagneticEnvironment(args, spec=spec, device=device, seed=21)
        runners.append(BatchedEpisodeRunner(env, copy.deepcopy(mac0), EpisodeReplayBuffer(args, device=device), args))
    for ep in range(3):
        info_s = runners[0].run(use_graph=False)
        info_g = runners[1].run(use_graph=True)
        assert not getattr(runners[1], "_graph_failed", False)
        for k, v in runners[0].traj.items():
            assert torch.equal(v, runners[1].traj[k]), (ep, k)
        assert info_s["episode_return"] == info_g["episode_return"]
    assert runners[0].t_env == runners[1].t_env and runners[0].mac._rng_step == runners[1].mac._rng_step


def check_fast_path_chains_hidden_state(device, lib, n_envs=10):
    """The cached-struct path of BatchedEpisodeRunner chains the recurrent state through the trajectory's h_t
    records (hidden_in = slot t - 1, only the last step writes mac.hidden_states); a twin runner that takes the
    general path (in-place mac.hidden_states every step) must produce the same episode, twice in a row."""
    import copy
    from macjd_b200.simulation.environment import ElectromagneticEnvironment
    from macjd_b200.simulation.scenario import hetero_spec
    from macjd_b200.core.mac import BasicMAC
    from macjd_b200.runners.episode_runner import BatchedEpisodeRunner
    args = rl_args(device)
    spec = hetero_spec(n_envs, seed=12, active=True, episode_limit=args.episode_limit)
    torch.manual_seed(2)
    mac0 = BasicMAC(24, args, _lib=lib)
    if args.use_cuda:
        mac0.cuda()
    runners = [BatchedEpisodeRunner(ElectromagneticEnvironment(args, spec=spec, device=device, seed=5, _lib=lib),
                                    copy.deepcopy(mac0), None, args) for _ in range(2)]
    T = args.episode_limit
    never = torch.ones(n_envs * args.n_agents, device=device)          # u >= epsilon: no exploration, general path
    for ep in range(2):
        runners[0].run(test_mode=True, store=False)
        runners[1].reset()
        for t in range(T):
            runners[1].step(t, test_mode=True, u_eps=never)
        for k, v in runners[0].traj.items():
            assert torch.equal(v, runners[1].traj[k]), (ep, k)
        for r in runners:
            assert torch.equal(r.mac.hidden_states.view(-1), r.traj["hidden_state"][T - 1].reshape(-1))


def check_main_loop(device, lib, tmp_path):
    """macjd_b200.main.run (main.py:72-289 on the batched path): counters with the reference's meaning,
    training starts after start_training_steps, greedy evaluation leaves the exploration schedule alone,
    the reference's console lines / TensorBoard tags, and checkpoints that load back."""
    from macjd_b200 import main as M
    from macjd_b200.simulation.scenario import hetero_spec
    n_envs, T = 6, 5
    args = rl_args(device, episode_limit=T, batch_size=4, buffer_size=8, total_env_steps=4 * n_envs * T,
                   start_training_steps=n_envs * T, train_interval=2, log_interval=10, log_interval_seconds=0,
                   save_model=True, save_model_dir=str(tmp_path / "models"), save_interval=2 * n_envs, test_name="t",
                   test_interval=2 * n_envs * T, test_nepisodes=7, device_request=device)
    lines, scalars = [], []
    writer = types.SimpleNamespace(add_scalar=lambda tag, v, step: scalars.append((tag, float(v), step)), close=lambda: None)
    out = M.run(args, spec=hetero_spec(n_envs, seed=3, active=True, episode_limit=T), writer=writer, log=lines.append, _lib=lib)
    assert out["episodes"] == 4 * n_envs and out["total_steps"] == 4 * n_envs * T
    # rollouts 2, 3 and 4 train; the reference's update-to-data ratio: one train step per train_interval
    # single-environment steps (main.py:216), i.e. n_envs * (T // train_interval) per rollout
    assert out["train_steps"] == 3 * n_envs * (T // 2)
    ref_ratio = (T // 2) / T                                     # reference: T // train_interval train steps per T env steps
    assert abs(out["train_steps"] / (3 * n_envs * T) - ref_ratio) < 1e-12
    assert out["buffer"].buffer_size == 8 and len(out["buffer"]) == 8
    assert out["runner"].t_env == 4 * n_envs * T                 # single-env steps; two evaluations did not advance the schedule
    # epsilon(total_steps) is the reference's schedule (action_selectors.py:30-32) evaluated on single-env steps
    last_tick = 4 * n_envs * T - n_envs                          # t_env of the last batched timestep that acted
    want_eps = max(args.epsilon_finish, args.epsilon_start - (args.epsilon_start - args.epsilon_finish) / args.epsilon_anneal_time * last_tick)
    assert abs(out["runner"].mac.action_selector.epsilon - want_eps) < 1e-9
    assert out["last_eval"]["n_episodes"] == 2 * n_envs and np.isfinite(out["last_eval"]["episode_return"])
    assert abs(out["last_logged"]["action_dist"].sum() - 1) < 1e-6 and np.isfinite(out["last_logged"]["avg_loss"])
    tags = {t for t, _, _ in scalars}
    assert {"Perf/Avg_Return", "Perf/Avg_Length", "Perf/Avg_Step_Reward", "Loss/train_avg", "Loss/train_episode_avg",
            "Params/Epsilon", "Params/Buffer_Size", "Stats/grad_norm", "QValues/eval_qtot_avg", "QValues/target_qtot_avg",
            "Rewards/r_d_avg", "Rewards/r_p_avg", "Rewards/r_j_avg", "Perf/Avg_Power", "ActionDist/Action_0",
            "Test/Avg_Return"} <= tags
    assert lines[0] == "Starting training..." and lines[-1] == "Training finished."
    assert any(l.startswith(f"Steps: {4 * n_envs * T}/{4 * n_envs * T} | Episodes: {4 * n_envs}") for l in lines)
    assert set(os.listdir(tmp_path / "models" / "t")) == {f"step_{k * n_envs * T}" for k in (2, 4)}
    d = tmp_path / "models" / "t" / f"step_{4 * n_envs * T}"
    assert sorted(os.listdir(d)) == ["agent.pth", "optimizer.pth", "qmix_net.pth"]
    sd = torch.load(d / "agent.pth", map_location="cpu")
    for k, v in out["learner"].mac.agent.state_dict().items():
        assert torch.equal(sd[k], v.cpu()), k
    with pytest.raises(FileNotFoundError):
        M.load_config("nope", str(tmp_path))
    cfg = M.default_config(lr=1e-4)
    assert cfg.lr == 1e-4 and cfg.batch_size == 32 and cfg.target_update_interval == 200


def check_fused_rollout_step(device, lib, n_envs=300, kind="c2", path=0, expect_fused=None):
    """macjd_rollout_step (one launch when the CTA-pair kernel can run the env step of its rows' envs itself)
    against macjd_agent_forward + macjd_env_step: whole episodes with device Philox draws on twin runners must be
    bit-identical in every trajectory tensor, reward part and step counter."""
    import copy
    from macjd_b200 import _native as N
    from macjd_b200.simulation.environment import ElectromagneticEnvironment
    from macjd_b200.simulation.scenario import hetero_spec, scaled_spec
    from macjd_b200.core.mac import BasicMAC
    from macjd_b200.runners.episode_runner import BatchedEpisodeRunner
    if kind == "c2":
        spec, dims = hetero_spec(n_envs, seed=31, active=True, episode_limit=5), dict()
        J, S, A = 2, 24, 5
    elif kind == "c3":
        J, R, K = 8, 16, 4
        spec = scaled_spec(n_envs, n_jammers=J, n_radars=R, n_targets=K, seed=32, episode_limit=5)
        S, A = R * 10 + 2 * J, 2 * R + 1
    else:                       # 3 jammers: a CTA's 64 rows are not whole envs -> two launches
        J, R, K = 3, 2, 2
        spec = scaled_spec(n_envs, n_jammers=J, n_radars=R, n_targets=K, seed=33, episode_limit=5)
        S, A = R * 10 + 2 * J, 2 * R + 1
    args = rl_args(device, episode_limit=5, n_agents=J, n_actions=A, state_shape=S, obs_shape=S, rnn_hidden_dim=128,
                   actor_hidden_dim=128, agent_kernel_path=path, epsilon_anneal_time=10 * n_envs)
    torch.manual_seed(4)
    mac0 = BasicMAC(S, args, _lib=lib)
    if args.use_cuda:
        mac0.cuda()
    runners = []
    # whole episode as one call (one launch when fused) / the same in two chunks / one fused call per step / two kernels per step
    for fused, whole in ((True, True), (True, "chunks"), (True, False), (False, False)):
        env = ElectromagneticEnvironment(args, spec=spec, device=device, seed=17, _lib=lib)
        r = BatchedEpisodeRunner(env, copy.deepcopy(mac0), None, args)
        r.fused_step, r.whole_episode_launch = fused, whole is True
        r.chunks = whole == "chunks"
        runners.append(r)
    if expect_fused is not None and lib is None:
        fn = N.get_lib().lib.macjd_rollout_fused_supported
        fn.restype = N.C.c_int
        got = fn(N.C.byref(runners[0].mac.agent.packed().cstruct()), N.C.byref(runners[0].env._ctab))
        assert bool(got) == expect_fused
    for ep in range(2):
        for r in runners:
            if r.chunks:
                r.reset()
                r.rollout(0, 2)
                r.rollout(2, args.episode_limit - 2)
            else:
                r.run(store=False)
        b = runners[-1]
        for i, a in enumerate(runners[:-1]):
            for k in a.traj:
                assert torch.equal(a.traj[k], b.traj[k]), (ep, i, k)
            assert torch.equal(a.r_parts, b.r_parts), (ep, i)
            assert torch.equal(a.env.step_count, b.env.step_count), (ep, i)
            assert torch.equal(a.mac.hidden_states, b.mac.hidden_states), (ep, i)
            assert a.t_env == b.t_env and a.mac._rng_step == b.mac._rng_step
            for k in ("pd", "detected", "tracking", "snr1", "jam_power", "pd_net", "reward64"):
                assert torch.equal(getattr(a.env, k), getattr(b.env, k)), (ep, i, k)
        assert bool(b.traj["terminated"][-1].all()) and not bool(b.traj["terminated"][0].any())
