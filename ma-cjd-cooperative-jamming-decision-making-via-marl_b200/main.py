"""Training loop on the batched, device-resident path (drop-in for main.py:72-289).

The reference's loop is  run one episode -> store -> (sample -> train) x episode_length // train_interval ->
log -> checkpoint  on one environment.  Here one iteration advances ``n_envs`` episodes together
(``BatchedEpisodeRunner``: two fused kernels per timestep, trajectories written straight into the replay
layout) and stores them in the HBM replay ring.

Schedules keep the reference's meaning, per unit of EXPERIENCE:
  * ``episode`` counts episodes and ``total_steps`` single-environment steps (main.py:196-198), so
    ``total_env_steps``, ``start_training_steps``, ``save_interval`` carry over;
  * the exploration schedule is annealed on ``runner.t_env`` = total single-environment steps, as in the
    reference (mac.py:96): a batched timestep advances it by ``n_envs``;
  * the update-to-data ratio is the reference's: one train step per ``train_interval`` environment steps
    (main.py:216: ``episode_length // train_interval`` per episode), i.e. ``n_envs * (T // train_interval)``
    train steps per rollout.  ``args.updates_per_env_step`` (float) overrides the ratio and
    ``args.train_steps_per_rollout`` (int) fixes the count -- at thousands of envs the reference's ratio makes
    the loop learner-bound by three orders of magnitude, and a user may not want that.
Console lines and TensorBoard tags are the reference's (main.py:232-277); checkpoints are
``learner.save_models`` directories (agent.pth / qmix_net.pth / optimizer.pth, interchangeable with the
reference's, main.py:280-286).

``pipeline=True``: rollout k runs on one stream with a frozen ACTING copy of the agent
while the learner trains on another stream from the ring as it stood after rollout k - 1; the copy is
refreshed (a device copy) and the ring slots are handed over between rollouts with events, statistics are
read back once per log line (``lazy_stats``), and no train step synchronises with the host.  The acting
policy therefore lags the learner by one rollout -- the usual actor / learner split; ``pipeline=False`` is
the reference's strict alternation (act with the latest weights, then train).  The pipelined loop also draws its
batches with the ring's O(batch) sampler (``args.replay_fast_sampling``, default on here, off everywhere else: same
distribution, different index stream than ``np.random.choice``'s full permutation, which at a 125 000-episode ring
costs the host four train steps per draw); ``replay_fast_sampling=False`` keeps the reference's draws.

Not in the reference: ``evaluate`` (greedy episodes with ``test_mode=True``; the reference's
``test_interval`` / ``test_nepisodes`` keys of config/default.yaml:81-83 are never read by its loop) and
``n_envs`` / ``spec`` for the batched environment.  The reference's command line is not rebuilt
(out of scope); ``tools/train_cli.py`` is a development convenience.

Everything on the device goes through the CUDA library; there is no CPU fallback behind this loop.
"""
from __future__ import annotations

import copy
import os
import time
from collections import deque
from datetime import datetime
from types import SimpleNamespace

import numpy as np
import torch
import yaml

from .core.mac import BasicMAC
from .core.qmix import QMixLearner, TRAINED_AGENT_KEYS
from .runners.episode_runner import BatchedEpisodeRunner
from .simulation.environment import ElectromagneticEnvironment
from .utils.replay_buffer import EpisodeReplayBuffer

# config/default.yaml of the reference, restated (values only; the file itself is not shipped)
DEFAULTS = dict(
    seed=42, use_cuda=True, device="cuda", batch_size=32, buffer_size=5000, lr=5e-6, gamma=0.99, grad_norm_clip=1.0,
    target_update_interval=200, start_training_steps=1000, train_interval=1, epsilon_start=1.0, epsilon_finish=0.05,
    epsilon_anneal_time=100000, rnn_hidden_dim=128, actor_hidden_dim=128, mixing_embed_dim=64, hyper_hidden_dim=128,
    log_interval=100, log_interval_seconds=10, save_model=True, save_model_dir="models/", save_interval=50,
    test_name="ma_cjd_test", total_env_steps=20000, test_interval=10000, test_nepisodes=20, env_config="simulation_config")


def load_config(config_name="default", config_dir="config"):
    """main.py:40-70: a YAML file as a namespace; FileNotFoundError / yaml.YAMLError propagate."""
    path = os.path.join(config_dir, f"{config_name}.yaml")
    with open(path, "r") as f:
        return SimpleNamespace(**yaml.safe_load(f))


def default_config(**overrides):
    """The reference's default hyper-parameters (config/default.yaml) with ``overrides`` applied."""
    return SimpleNamespace(**{**DEFAULTS, **overrides})


def evaluate(runner, n_episodes):
    """Greedy evaluation (test_mode=True: no exploration, nothing stored, epsilon schedule untouched):
    at least ``n_episodes`` episodes in batches of ``runner.n_envs``; averages of the runner's statistics."""
    t_env = runner.t_env
    infos = []
    for _ in range(max(1, -(-int(n_episodes) // runner.n_envs))):
        infos.append(runner.run(test_mode=True, store=False))
    runner.t_env = t_env                               # evaluation does not advance the exploration schedule
    keys = ("episode_return", "avg_step_reward", "avg_r_d", "avg_r_p", "avg_r_j", "avg_power_overall")
    out = {k: float(np.mean([i[k] for i in infos])) for k in keys}
    out["action_distribution"] = np.mean(np.stack([i["action_distribution"] for i in infos]), axis=0)
    out["n_episodes"] = len(infos) * runner.n_envs
    return out


def _mean(q):
    return float(np.mean(q)) if len(q) else 0.0


def train_steps_for_rollout(args, n_envs, episode_steps):
    """How many train steps follow a rollout of ``n_envs`` episodes of ``episode_steps`` steps: the
    reference's update-to-data ratio (main.py:216) unless the config overrides it."""
    fixed = getattr(args, "train_steps_per_rollout", None)
    if fixed is not None:
        return max(0, int(fixed))
    ratio = getattr(args, "updates_per_env_step", None)
    if ratio is not None:
        return max(0, int(round(float(ratio) * n_envs * episode_steps)))
    return n_envs * (episode_steps // max(1, args.train_interval))


def run(args, *, n_envs=None, spec=None, sim_config_path=None, writer=None, log=print, use_graph=False, pipeline=False,
        _lib=None):
    """main.py:72-289 on the batched path.  ``args`` is the reference's config namespace (missing keys take
    ``DEFAULTS``); ``n_envs`` (or ``args.n_envs``) episodes advance per iteration, from ``sim_config_path``
    (the reference's scenario YAML replicated) or a ``ScenarioSpec``.  ``writer``: a TensorBoard
    ``SummaryWriter``-like object, ``None`` = create one under logs/<test_name>/ as the reference does,
    ``False`` = no TensorBoard.  ``pipeline``: see the module docstring.  Returns the final counters and the
    last logged averages.  (``_lib``: the host-emulation build of the kernels, tests only.)"""
    for k, v in DEFAULTS.items():
        if not hasattr(args, k):
            setattr(args, k, v)
    # main.py:84-121: device choice and seeds -- no CPU fallback here: the kernels are the product
    requested = str(getattr(args, "device_request", getattr(args, "device", "cuda"))).lower()
    if _lib is not None:
        args.device, args.use_cuda = "cpu", False                 # kernels compiled for the host (tests/emul)
        pipeline = False
    else:
        if not requested.startswith("cuda"):
            raise RuntimeError("macjd_b200.main.run needs a CUDA device (no CPU path behind it)")
        if not torch.cuda.is_available():
            raise RuntimeError("CUDA requested but not available")
        args.device = requested if ":" in requested else f"cuda:{torch.cuda.current_device()}"
        args.use_cuda = True
    np.random.seed(args.seed)
    torch.manual_seed(args.seed)
    if args.use_cuda:
        torch.cuda.manual_seed(args.seed)

    test_name = getattr(args, "test_name", "ma_cjd_test")
    own_writer = False
    if writer is None:
        from torch.utils.tensorboard import SummaryWriter
        log_dir = os.path.join("logs", test_name, f"run_{datetime.now().strftime('%Y%m%d_%H%M%S')}")
        writer, own_writer = SummaryWriter(log_dir=log_dir), True
        log(f"TensorBoard logs will be saved to: {log_dir}")
    elif writer is False:
        writer = None

    # main.py:133-160: components
    if n_envs is None and getattr(args, "n_envs", None) is None and spec is None:
        # a rollout should be a small part of the step budget (the loop checks the budget between rollouts)
        n_envs = int(min(1024, max(1, args.total_env_steps // (8 * max(1, getattr(args, "episode_limit", 100))))))
    n_envs = int(n_envs or getattr(args, "n_envs", None) or 1)
    if spec is not None:
        env = ElectromagneticEnvironment(args, spec=spec, device=args.device, seed=args.seed, _lib=_lib)
    else:
        path = sim_config_path or os.path.join("config", f"{args.env_config}.yaml")
        env = ElectromagneticEnvironment(args, sim_config_path=path, n_envs=n_envs, device=args.device, seed=args.seed, _lib=_lib)
    n_envs = env.n_envs
    env_info = env.get_env_info()
    args.n_agents, args.n_actions = env_info["n_agents"], env_info["n_actions"]
    args.state_shape = env_info["state_shape"]
    args.obs_shape = env_info.get("obs_shape", args.state_shape)
    args.episode_limit = env_info["episode_limit"]
    args.env_info = env_info
    args.buffer_size = max(int(args.buffer_size), n_envs)        # the ring must hold one rollout
    mac = BasicMAC(input_shape=args.obs_shape, args=args, _lib=_lib)
    if args.use_cuda:
        mac.cuda()
    # The ring drops `obs` only when the ENVIRONMENT says every agent observes the replicated global state
    # (environment.py:512-522; ElectromagneticEnvironment.obs_is_replicated_state) or the config asks for it:
    # equal widths alone do not prove equal contents.
    shared = getattr(args, "replay_shared_obs", None)
    if shared is None:
        shared = bool(getattr(env, "obs_is_replicated_state", False)) and args.obs_shape == args.state_shape
    # the pipelined loop draws its batches with the O(batch) sampler unless told otherwise: the reference's draw permutes
    # the whole ring population per batch (2.7 ms at 125 000 episodes; utils/replay_buffer.py header)
    fast = getattr(args, "replay_fast_sampling", None)
    fast = (bool(pipeline) and args.use_cuda) if fast is None else bool(fast)
    buffer = EpisodeReplayBuffer(args=args, device=args.device, _lib=_lib, shared_obs=bool(shared), fast_sampling=fast)
    learner = QMixLearner(mac, args=args, _lib=_lib)
    pipeline = bool(pipeline) and args.use_cuda
    # pipelined: the rollout acts with a frozen copy of the agent, refreshed between rollouts
    act_mac = copy.deepcopy(mac) if pipeline else mac
    runner = BatchedEpisodeRunner(env=env, mac=act_mac, buffer=buffer, args=args)
    if pipeline:
        dev = torch.device(args.device)
        roll_stream, learn_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        ev_weights, ev_gathered, ev_stored, ev_copied = (torch.cuda.Event() for _ in range(4))
        main_stream = torch.cuda.current_stream(dev)
        roll_stream.wait_stream(main_stream)
        learn_stream.wait_stream(main_stream)
        ev_weights.record(learn_stream)
        ev_gathered.record(learn_stream)

    # MACJD_LOOP_TIMING=1: device time of each rollout / train phase / store of the pipelined loop (CUDA events on their
    # streams; returned as out["phase_ms"]: medians over the iterations)
    phase_events = {} if pipeline and os.environ.get("MACJD_LOOP_TIMING", "0") != "0" else None

    def stamp(name, stream):
        if phase_events is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(stream)
            phase_events.setdefault(name, []).append(e)

    def sync_acting_copy():
        """Acting copy <- learner's agent: only fc2_q_head is ever trained (core/qmix.py:178), so four raw
        device copies and the Q-head re-pack (no full re-pack, no version bump on the acting parameters)."""
        src = dict(mac.agent.named_parameters())
        with torch.no_grad():
            for k, p_ in act_mac.agent.named_parameters():
                if k in TRAINED_AGENT_KEYS:
                    p_.data.copy_(src[k].data)
        act_mac.agent.packed()
        act_mac.agent.packed_qhead()

    start_time = last_log_time = time.time()
    episode = total_steps = train_steps = 0
    last_test_steps = 0
    n_train_nominal = train_steps_for_rollout(args, n_envs, args.episode_limit)
    stats = {k: deque(maxlen=args.log_interval) for k in
             ("episode_return", "episode_length", "avg_step_reward", "reward_r_d", "reward_r_p", "reward_r_j",
              "avg_power", "action_dist")}
    stat_keys = ("loss", "grad_norm", "eval_qtot_avg", "target_qtot_avg")
    stats.update({k: deque(maxlen=args.log_interval * max(1, n_train_nominal)) for k in stat_keys})
    pending_stats = []                 # lazily read train statistics: (device tensor [4], total_steps at that time)
    last_logged, last_eval = {}, None

    def scalar(tag, value, step):
        if writer is not None:
            writer.add_scalar(tag, value, step)

    def flush_train_stats():
        """One read-back for every train step since the last flush (pipelined mode)."""
        if not pending_stats:
            return
        rows = torch.stack([t for t, _ in pending_stats]).tolist()
        by_step = {}
        for (_, at), r in zip(pending_stats, rows):
            for k, v in zip(stat_keys, r):
                stats[k].append(v)
            by_step.setdefault(at, []).append(r[0])
        for at, losses in by_step.items():
            scalar("Loss/train_episode_avg", float(np.mean(losses)), at)
        pending_stats.clear()

    def train_phase(n_steps, at_steps):
        count, loss_sum = 0, 0.0
        for _ in range(n_steps):
            if pipeline:
                # sample + train as one call: on one GPU the whole step is a replayed CUDA graph (QMixLearner.train_sampled)
                ts = learner.train_sampled(buffer, args.batch_size, {"total_steps": at_steps})
                if ts is None:
                    continue
                pending_stats.append((ts["stats_tensor"], at_steps))
                count += 1
                continue
            batch = buffer.sample(args.batch_size, time_major=True)
            if batch is None:
                continue
            ts = learner.train(batch, {"total_steps": at_steps})
            for k in stat_keys:
                stats[k].append(ts[k])
            loss_sum += ts["loss"]
            count += 1
        if count and not pipeline:
            scalar("Loss/train_episode_avg", loss_sum / count, at_steps)
        return count

    log("Starting training...")
    while total_steps < args.total_env_steps:
        episodes_before = episode
        will_train = buffer.current_size >= args.batch_size and total_steps > args.start_training_steps
        if pipeline:
            # ---- rollout k on roll_stream with the acting copy, train on learn_stream from the ring as of k - 1
            with torch.cuda.stream(roll_stream):
                roll_stream.wait_event(ev_weights)              # the learner's weights of the previous phase are final
                sync_acting_copy()
                ev_copied.record(roll_stream)
            # (A/B on one B200, bench C5, four runs with temporary switches: issuing the learner's phase before the
            # rollout, or the rollout as ONE launch per episode, is slower -- 41.6 M / 38.6 M env-agent steps/s against
            # 44.5 M, both together 32.9 M.  Likely cause: the two streams compete for the SMs' shared memory, and a 2 ms
            # launch on 128 SMs holds up the learner's chain of short kernels for its whole length, while per-step
            # launches interleave with it.)
            with torch.cuda.stream(roll_stream):
                runner.reset()
                stamp("rollout", roll_stream)
                # use_graph: the episode's launches replayed as one CUDA graph -- the same per-step launches on the device,
                # but the host issues them in ~0.1 ms instead of ~2 ms and gets to the learner's phase that much earlier
                graphed = False
                if use_graph and not getattr(runner, "_graph_failed", False):
                    try:
                        runner._run_graph(False)
                        graphed = True
                    except RuntimeError:                        # capture not possible here: keep the per-step launches
                        runner._graph_failed = True
                        runner.reset()
                if not graphed:
                    for t in range(args.episode_limit):
                        runner.step(t)
                stamp("rollout", roll_stream)
            if will_train:
                with torch.cuda.stream(learn_stream):
                    learn_stream.wait_event(ev_stored)          # ring slots written by the previous store
                    learn_stream.wait_event(ev_copied)          # the acting copy has read the weights this phase updates
                    stamp("train", learn_stream)
                    train_steps += train_phase(train_steps_for_rollout(args, n_envs, args.episode_limit), total_steps)
                    stamp("train", learn_stream)
                    ev_gathered.record(learn_stream)            # (every gather of this phase is enqueued before this)
                    ev_weights.record(learn_stream)
            with torch.cuda.stream(roll_stream):
                roll_stream.wait_event(ev_gathered)             # the store may overwrite slots the learner sampled
                stamp("store", roll_stream)
                run_info = runner.finish_run(store=True)
                stamp("store", roll_stream)
                ev_stored.record(roll_stream)
        else:
            # ---- one rollout: n_envs episodes (main.py:193-198)
            run_info = runner.run(test_mode=False, use_graph=use_graph)
        episode += n_envs
        current_episode_steps = run_info["episode_length"]
        total_steps += current_episode_steps * n_envs
        stats["episode_return"].append(run_info["episode_return"])
        stats["episode_length"].append(current_episode_steps)
        stats["avg_step_reward"].append(run_info.get("avg_step_reward", 0))
        stats["reward_r_d"].append(run_info.get("avg_r_d", 0))
        stats["reward_r_p"].append(run_info.get("avg_r_p", 0))
        stats["reward_r_j"].append(run_info.get("avg_r_j", 0))
        stats["avg_power"].append(run_info.get("avg_power_overall", 0))
        stats["action_dist"].append(run_info["action_distribution"])

        # ---- learner (main.py:212-228), strict alternation
        if not pipeline and buffer.current_size >= args.batch_size and total_steps > args.start_training_steps:
            train_steps += train_phase(train_steps_for_rollout(args, n_envs, current_episode_steps), total_steps)

        # ---- greedy evaluation (not in the reference's loop; its config names the two keys)
        if getattr(args, "test_nepisodes", 0) and getattr(args, "test_interval", 0) and \
                total_steps - last_test_steps >= args.test_interval:
            last_test_steps = total_steps
            if pipeline:
                torch.cuda.synchronize(dev)
                sync_acting_copy()
            last_eval = evaluate(runner, args.test_nepisodes)
            if pipeline:                                          # the evaluation ran on the caller's stream
                roll_stream.wait_stream(main_stream)
            log(f"  Test ({last_eval['n_episodes']} greedy eps): Return {last_eval['episode_return']:.2f} | "
                f"r_d/r_p/r_j {last_eval['avg_r_d']:.4f} / {last_eval['avg_r_p']:.4f} / {last_eval['avg_r_j']:.4f}")
            scalar("Test/Avg_Return", last_eval["episode_return"], total_steps)
            scalar("Test/Avg_Step_Reward", last_eval["avg_step_reward"], total_steps)
            scalar("Test/Avg_Power", last_eval["avg_power_overall"], total_steps)

        # ---- logging (main.py:231-277): same lines, same tags
        now = time.time()
        if now - last_log_time >= args.log_interval_seconds or total_steps >= args.total_env_steps:
            flush_train_stats()
            dist = np.mean(np.array(stats["action_dist"]), axis=0) if stats["action_dist"] else np.zeros(args.n_actions)
            L = last_logged = {
                "avg_return": _mean(stats["episode_return"]), "avg_length": _mean(stats["episode_length"]),
                "avg_step_reward": _mean(stats["avg_step_reward"]), "avg_loss": _mean(stats["loss"]),
                "avg_grad_norm": _mean(stats["grad_norm"]), "avg_eval_qtot": _mean(stats["eval_qtot_avg"]),
                "avg_target_qtot": _mean(stats["target_qtot_avg"]), "avg_r_d": _mean(stats["reward_r_d"]),
                "avg_r_p": _mean(stats["reward_r_p"]), "avg_r_j": _mean(stats["reward_r_j"]),
                "avg_power": _mean(stats["avg_power"]), "action_dist": dist}
            log(f"Steps: {total_steps}/{args.total_env_steps} | Episodes: {episode} | Time: {now - start_time:.2f}s")
            log(f"  Avg Return (last {len(stats['episode_return'])} eps): {L['avg_return']:.2f} | "
                f"Avg Length: {L['avg_length']:.1f} | Avg Loss: {L['avg_loss']:.4f}")
            log(f"  Avg Step Reward (last {len(stats['avg_step_reward'])} eps): {L['avg_step_reward']:.4f}")
            log(f"  Avg Rewards (r_d/r_p/r_j): {L['avg_r_d']:.4f} / {L['avg_r_p']:.4f} / {L['avg_r_j']:.4f}")
            log(f"  Avg QTot (Eval/Target): {L['avg_eval_qtot']:.4f} / {L['avg_target_qtot']:.4f} | "
                f"Avg Grad Norm: {L['avg_grad_norm']:.4f}")
            log(f"  Avg Power: {L['avg_power']:.3f} | Action Dist: [{' / '.join(f'{p:.2f}' for p in dist)}] "
                f"(0=Idle, 1=S0, 2=D0, ...)")
            log(f"  Buffer Size: {len(buffer)}")
            log(f"  Epsilon: {act_mac.action_selector.epsilon:.3f}")
            for tag, key in (("Perf/Avg_Return", "avg_return"), ("Perf/Avg_Length", "avg_length"),
                             ("Perf/Avg_Step_Reward", "avg_step_reward"), ("Loss/train_avg", "avg_loss"),
                             ("Stats/grad_norm", "avg_grad_norm"), ("QValues/eval_qtot_avg", "avg_eval_qtot"),
                             ("QValues/target_qtot_avg", "avg_target_qtot"), ("Rewards/r_d_avg", "avg_r_d"),
                             ("Rewards/r_p_avg", "avg_r_p"), ("Rewards/r_j_avg", "avg_r_j"), ("Perf/Avg_Power", "avg_power")):
                scalar(tag, L[key], total_steps)
            scalar("Params/Epsilon", act_mac.action_selector.epsilon, total_steps)
            scalar("Params/Buffer_Size", len(buffer), total_steps)
            for a, p in enumerate(dist):
                scalar(f"ActionDist/Action_{a}", p, total_steps)
            last_log_time = now

        # ---- checkpoint (main.py:280-286): whenever a multiple of save_interval episodes was passed
        crossed = episode // max(1, args.save_interval) > episodes_before // max(1, args.save_interval)
        if args.save_model and (crossed or total_steps >= args.total_env_steps) and total_steps > args.start_training_steps:
            save_dir = os.path.join(args.save_model_dir, test_name, f"step_{total_steps}")
            os.makedirs(save_dir, exist_ok=True)
            log(f"Saving model to {save_dir}")
            if pipeline:
                learn_stream.synchronize()
            learner.save_models(save_dir)

    if pipeline:
        torch.cuda.synchronize(dev)
        main_stream.wait_stream(roll_stream)
        main_stream.wait_stream(learn_stream)
        flush_train_stats()
    runner.close_env()
    if own_writer:
        writer.close()
    log("Training finished.")
    phase_ms = None
    if phase_events:
        phase_ms = {k: float(np.median([a.elapsed_time(b) for a, b in zip(v[0::2], v[1::2])])) for k, v in phase_events.items()}
    return {"episodes": episode, "total_steps": total_steps, "train_steps": train_steps, "time_s": time.time() - start_time,
            "phase_ms": phase_ms,
            "last_logged": last_logged, "last_eval": last_eval, "learner": learner, "runner": runner, "buffer": buffer,
            "pipeline": pipeline}
