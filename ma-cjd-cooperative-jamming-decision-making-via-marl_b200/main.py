"""Training loop on the batched, device-resident path (drop-in for main.py:72-289).

The reference's loop is  run one episode -> store -> (sample -> train) x episode_length // train_interval ->
log -> checkpoint  on one environment.  Here one iteration advances ``n_envs`` episodes together
(``BatchedEpisodeRunner``: two fused kernels per timestep, trajectories written straight into the replay
layout) and stores them in the HBM replay ring; the learner part of the iteration is the reference's:
``episode_length // train_interval`` train steps on uniformly sampled batches once the buffer holds
``batch_size`` episodes and ``total_steps > start_training_steps``.  Counters keep the reference's meaning --
``episode`` counts episodes, ``total_steps`` counts single-environment steps (main.py:196-198) -- so
schedules written for the reference (``total_env_steps``, ``start_training_steps``, ``save_interval``) carry
over; the exploration schedule advances one tick per batched timestep, ``BatchedEpisodeRunner.t_env``
(mac.py:96 reads the runner's counter there too).  Console lines and TensorBoard tags are the reference's
(main.py:232-277); checkpoints are ``learner.save_models`` directories (agent.pth / qmix_net.pth /
optimizer.pth, interchangeable with the reference's, main.py:280-286).

Not in the reference: ``evaluate`` (greedy episodes with ``test_mode=True``; the reference's
``test_interval`` / ``test_nepisodes`` keys of config/default.yaml:81-83 are never read by its loop) and
``n_envs`` / ``spec`` for the batched environment.

Everything on the device goes through the CUDA library; there is no CPU fallback behind this loop.
"""
from __future__ import annotations

import os
import time
from collections import deque
from datetime import datetime
from types import SimpleNamespace

import numpy as np
import torch
import yaml

from .core.mac import BasicMAC
from .core.qmix import QMixLearner
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


def run(args, *, n_envs=None, spec=None, sim_config_path=None, writer=None, log=print, use_graph=False, _lib=None):
    """main.py:72-289 on the batched path.  ``args`` is the reference's config namespace (missing keys take
    ``DEFAULTS``); ``n_envs`` (or ``args.n_envs``) episodes advance per iteration, from ``sim_config_path``
    (the reference's scenario YAML replicated) or a ``ScenarioSpec``.  ``writer``: a TensorBoard
    ``SummaryWriter``-like object, ``None`` = create one under logs/<test_name>/ as the reference does,
    ``False`` = no TensorBoard.  Returns the final counters and the last logged averages.
    (``_lib``: the host-emulation build of the kernels, tests only.)"""
    for k, v in DEFAULTS.items():
        if not hasattr(args, k):
            setattr(args, k, v)
    # main.py:84-121: device choice and seeds -- no CPU fallback here: the kernels are the product
    requested = str(getattr(args, "device_request", getattr(args, "device", "cuda"))).lower()
    if _lib is not None:
        args.device, args.use_cuda = "cpu", False                 # kernels compiled for the host (tests/emul)
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
    n_envs = int(n_envs or getattr(args, "n_envs", 1024))
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
    # every agent observes the global state (environment.py:512-522): the ring keeps the state only
    buffer = EpisodeReplayBuffer(args=args, device=args.device, _lib=_lib,
                                 shared_obs=getattr(args, "replay_shared_obs", args.obs_shape == args.state_shape))
    learner = QMixLearner(mac, args=args, _lib=_lib)
    runner = BatchedEpisodeRunner(env=env, mac=mac, buffer=buffer, args=args)

    start_time = last_log_time = time.time()
    episode = total_steps = train_steps = 0
    last_test_steps = 0
    n_train = args.episode_limit // max(1, args.train_interval)
    stats = {k: deque(maxlen=args.log_interval) for k in
             ("episode_return", "episode_length", "avg_step_reward", "reward_r_d", "reward_r_p", "reward_r_j",
              "avg_power", "action_dist")}
    stats.update({k: deque(maxlen=args.log_interval * max(1, n_train)) for k in
                  ("loss", "grad_norm", "eval_qtot_avg", "target_qtot_avg")})
    last_logged, last_eval = {}, None

    def scalar(tag, value, step):
        if writer is not None:
            writer.add_scalar(tag, value, step)

    log("Starting training...")
    while total_steps < args.total_env_steps:
        # ---- one rollout: n_envs episodes (main.py:193-198)
        run_info = runner.run(test_mode=False, use_graph=use_graph)
        episodes_before = episode
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

        # ---- learner (main.py:212-228)
        if buffer.current_size >= args.batch_size and total_steps > args.start_training_steps:
            loss_sum, count = 0.0, 0
            for _ in range(current_episode_steps // max(1, args.train_interval)):
                batch = buffer.sample(args.batch_size, time_major=True)
                if batch is None:
                    continue
                ts = learner.train(batch, {"total_steps": total_steps})
                for k in ("loss", "grad_norm", "eval_qtot_avg", "target_qtot_avg"):
                    stats[k].append(ts[k])
                loss_sum += ts["loss"]
                count += 1
            train_steps += count
            if count:
                scalar("Loss/train_episode_avg", loss_sum / count, total_steps)

        # ---- greedy evaluation (not in the reference's loop; its config names the two keys)
        if getattr(args, "test_nepisodes", 0) and getattr(args, "test_interval", 0) and \
                total_steps - last_test_steps >= args.test_interval:
            last_test_steps = total_steps
            last_eval = evaluate(runner, args.test_nepisodes)
            log(f"  Test ({last_eval['n_episodes']} greedy eps): Return {last_eval['episode_return']:.2f} | "
                f"r_d/r_p/r_j {last_eval['avg_r_d']:.4f} / {last_eval['avg_r_p']:.4f} / {last_eval['avg_r_j']:.4f}")
            scalar("Test/Avg_Return", last_eval["episode_return"], total_steps)
            scalar("Test/Avg_Step_Reward", last_eval["avg_step_reward"], total_steps)
            scalar("Test/Avg_Power", last_eval["avg_power_overall"], total_steps)

        # ---- logging (main.py:231-277): same lines, same tags
        now = time.time()
        if now - last_log_time >= args.log_interval_seconds or total_steps >= args.total_env_steps:
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
            log(f"  Epsilon: {mac.action_selector.epsilon:.3f}")
            for tag, key in (("Perf/Avg_Return", "avg_return"), ("Perf/Avg_Length", "avg_length"),
                             ("Perf/Avg_Step_Reward", "avg_step_reward"), ("Loss/train_avg", "avg_loss"),
                             ("Stats/grad_norm", "avg_grad_norm"), ("QValues/eval_qtot_avg", "avg_eval_qtot"),
                             ("QValues/target_qtot_avg", "avg_target_qtot"), ("Rewards/r_d_avg", "avg_r_d"),
                             ("Rewards/r_p_avg", "avg_r_p"), ("Rewards/r_j_avg", "avg_r_j"), ("Perf/Avg_Power", "avg_power")):
                scalar(tag, L[key], total_steps)
            scalar("Params/Epsilon", mac.action_selector.epsilon, total_steps)
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
            learner.save_models(save_dir)

    runner.close_env()
    if own_writer:
        writer.close()
    log("Training finished.")
    return {"episodes": episode, "total_steps": total_steps, "train_steps": train_steps, "time_s": time.time() - start_time,
            "last_logged": last_logged, "last_eval": last_eval, "learner": learner, "runner": runner, "buffer": buffer}


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description="QMix / MP-DQN training on the batched device-resident path")
    ap.add_argument("--config", default=None, help="name of a YAML file under --config-dir (reference format); default: built-in defaults")
    ap.add_argument("--config-dir", default="config")
    ap.add_argument("--sim-config", default=None, help="scenario YAML (reference format); default: the reference's default scenario")
    ap.add_argument("--n-envs", type=int, default=1024)
    ap.add_argument("--total-env-steps", type=int, default=None)
    ap.add_argument("--no-tensorboard", action="store_true")
    a = ap.parse_args()
    cfg = load_config(a.config, a.config_dir) if a.config else default_config()
    if a.total_env_steps is not None:
        cfg.total_env_steps = a.total_env_steps
    if a.sim_config is None:
        from .simulation.scenario import default_spec
        run(cfg, spec=default_spec(a.n_envs), writer=False if a.no_tensorboard else None)
    else:
        run(cfg, n_envs=a.n_envs, sim_config_path=a.sim_config, writer=False if a.no_tensorboard else None)
