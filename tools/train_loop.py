#!/usr/bin/env python
"""End-to-end training loop on the device-resident path (BASELINE config 5 in miniature):
rollout of n_envs episodes -> store in the HBM replay ring -> sample -> train -> target sync.
Reports sustained env-agent steps/s and train samples/s.  Works under torchrun (one rank per GPU,
data-parallel learner).

  python tools/train_loop.py [--n-envs 4096] [--iters 5] [--train-steps 20] [--buffer 16384]
"""
import argparse
import json
import os
import sys
import time
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-envs", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--train-steps", type=int, default=20)
    ap.add_argument("--buffer", type=int, default=16384)
    ap.add_argument("--batch", type=int, default=32)
    a = ap.parse_args()
    import torch.distributed as dist
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = f"cuda:{local}"
    from macjd_b200.simulation.environment import ElectromagneticEnvironment
    from macjd_b200.simulation.scenario import default_spec
    from macjd_b200.core.mac import BasicMAC
    from macjd_b200.core.qmix import QMixLearner
    from macjd_b200.utils.replay_buffer import EpisodeReplayBuffer
    from macjd_b200.runners.episode_runner import BatchedEpisodeRunner
    rl = types.SimpleNamespace(
        n_agents=2, n_actions=5, state_shape=24, obs_shape=24, rnn_hidden_dim=128, actor_hidden_dim=128,
        mixing_embed_dim=64, hyper_hidden_dim=128, epsilon_start=1.0, epsilon_finish=0.05, epsilon_anneal_time=100000,
        gamma=0.99, lr=5e-6, grad_norm_clip=1.0, target_update_interval=200, use_cuda=True, device=dev,
        batch_size=a.batch, buffer_size=a.buffer, episode_limit=100, seed=42, data_parallel=True, agent_kernel_path=0)
    torch.manual_seed(42)
    env = ElectromagneticEnvironment(rl, spec=default_spec(a.n_envs), device=dev, seed=7 + rank)
    mac = BasicMAC(24, rl); mac.cuda()
    buf = EpisodeReplayBuffer(rl, device=dev)
    runner = BatchedEpisodeRunner(env, mac, buf, rl)
    learner = QMixLearner(mac, rl)
    np.random.seed(rank)
    runner.run()                                   # warm-up
    learner.train(buf.sample(a.batch, time_major=True), {})
    torch.cuda.synchronize()
    t_roll = t_train = 0.0
    info = stats = None
    for it in range(a.iters):
        t0 = time.perf_counter()
        info = runner.run()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        for _ in range(a.train_steps):
            stats = learner.train(buf.sample(a.batch, time_major=True), {})
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        t_roll += t1 - t0
        t_train += t2 - t1
    if rank == 0:
        steps = a.iters * a.n_envs * 100 * 2 * world
        print(json.dumps({
            "n_gpus": world, "rollout_env_agent_steps_per_sec": steps / t_roll,
            "rollout_ms_per_timestep": t_roll / (a.iters * 100) * 1e3,
            "train_episodes_per_sec": a.iters * a.train_steps * a.batch * world / t_train,
            "train_ms_per_step": t_train / (a.iters * a.train_steps) * 1e3,
            "loop_env_agent_steps_per_sec": steps / (t_roll + t_train),
            "replay_episodes": len(buf), "replay_gib_per_gpu": buf.bytes_per_episode() * len(buf) / 2**30,
            "last_episode_return": info["episode_return"], "last_loss": stats["loss"], "epsilon": mac.action_selector.epsilon}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
