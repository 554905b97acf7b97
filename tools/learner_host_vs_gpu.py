"""Is the QMix train step at the bench shape (B = 32 x T = 100) bound by the host or by the GPU?  Host issue time of
asynchronous train steps (lazy statistics, no read-back) against their GPU time, and a cProfile of the host side.
   python tools/learner_host_vs_gpu.py [n_steps]"""
import cProfile
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench   # noqa: E402
from macjd_b200.core.mac import BasicMAC   # noqa: E402
from macjd_b200.core.qmix import QMixLearner   # noqa: E402
from macjd_b200.runners.episode_runner import BatchedEpisodeRunner   # noqa: E402
from macjd_b200.simulation.environment import ElectromagneticEnvironment   # noqa: E402
from macjd_b200.simulation.scenario import default_spec   # noqa: E402
from macjd_b200.utils.replay_buffer import EpisodeReplayBuffer   # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 50
dev, n_envs = "cuda:0", 256
rl = bench.rl_args(dev, n_envs)
env = ElectromagneticEnvironment(rl, spec=default_spec(n_envs), device=dev, seed=1)
mac = BasicMAC(bench.OBS, rl)
mac.cuda()
buf = EpisodeReplayBuffer(rl, device=dev)
runner = BatchedEpisodeRunner(env, mac, buf, rl)
learner = QMixLearner(mac, rl)
runner.run()
np.random.seed(1)


def step():
    return learner.train(buf.sample(bench.LEARNER_B, time_major=True), {}, lazy_stats=True, check_actions=False)


for _ in range(5):
    step()
torch.cuda.synchronize()
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(N):
        step()
    e1.record()
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print(f"{N} async train steps: host issue {t_issue / N * 1e3:.3f} ms / step, GPU events {e0.elapsed_time(e1) / N:.3f} ms / step, "
          f"wall incl. drain {t_all / N * 1e3:.3f} ms / step")
# one step at a time: GPU time of an isolated step (the host runs ahead of nothing)
ts = []
for _ in range(10):
    b = buf.sample(bench.LEARNER_B, time_major=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    learner.train(b, {}, lazy_stats=True, check_actions=False)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("isolated train step (after a sync), GPU events ms:", " ".join(f"{x:.3f}" for x in ts))
# the call the training loop makes: sample + train as one replayed CUDA graph
for _ in range(4):
    learner.train_sampled(buf, bench.LEARNER_B, {})
torch.cuda.synchronize()
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(N):
        learner.train_sampled(buf, bench.LEARNER_B, {})
    e1.record()
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(f"{N} train_sampled steps (graph): host issue {t_issue / N * 1e3:.3f} ms / step, GPU events {e0.elapsed_time(e1) / N:.3f} ms / step")

pr = cProfile.Profile()
pr.enable()
for _ in range(N):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)

