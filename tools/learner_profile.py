"""Kernel / CPU breakdown of one QMixLearner.train step at the bench shape (B=32, T=100)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
from macjd_b200.simulation.environment import ElectromagneticEnvironment
from macjd_b200.simulation.scenario import default_spec
from macjd_b200.core.mac import BasicMAC
from macjd_b200.core.qmix import QMixLearner
from macjd_b200.utils.replay_buffer import EpisodeReplayBuffer
from macjd_b200.runners.episode_runner import BatchedEpisodeRunner

dev = "cuda:0"
n_envs = 256
rl = bench.rl_args(dev, n_envs)
env = ElectromagneticEnvironment(rl, spec=default_spec(n_envs), device=dev, seed=1)
mac = BasicMAC(bench.OBS, rl); mac.cuda()
buf = EpisodeReplayBuffer(rl, device=dev)
runner = BatchedEpisodeRunner(env, mac, buf, rl)
learner = QMixLearner(mac, rl)
runner.run()
np.random.seed(1)
for _ in range(5):
    learner.train(buf.sample(bench.LEARNER_B, time_major=True), {})
torch.cuda.synchronize()
t0 = time.perf_counter()
N = 20
for _ in range(N):
    learner.train(buf.sample(bench.LEARNER_B, time_major=True), {})
torch.cuda.synchronize()
print(f"wall per train step: {(time.perf_counter() - t0) / N * 1e3:.3f} ms")
t0 = time.perf_counter()
for _ in range(N):
    b = buf.sample(bench.LEARNER_B, time_major=True)
torch.cuda.synchronize()
print(f"  of which sample(): {(time.perf_counter() - t0) / N * 1e3:.3f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        learner.train(buf.sample(bench.LEARNER_B, time_major=True), {})
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
