"""Host cost of one rollout step (two ctypes launches) vs its GPU time: is runner.run() launch-bound?"""
import sys, time, cProfile, pstats
import torch
sys.path.insert(0, ".")
import bench
from macjd_b200.simulation.environment import ElectromagneticEnvironment
from macjd_b200.simulation.scenario import default_spec
from macjd_b200.core.mac import BasicMAC
from macjd_b200.utils.replay_buffer import EpisodeReplayBuffer
from macjd_b200.runners.episode_runner import BatchedEpisodeRunner
dev = "cuda:0"; n_envs = 4096
rl = bench.rl_args(dev, n_envs)
env = ElectromagneticEnvironment(rl, spec=default_spec(n_envs), device=dev, seed=1)
mac = BasicMAC(bench.OBS, rl); mac.cuda()
buf = EpisodeReplayBuffer(rl, device=dev)
runner = BatchedEpisodeRunner(env, mac, buf, rl)
runner.run(); runner.reset()
torch.cuda.synchronize()
t0 = time.perf_counter()
for t in range(100): runner.step(t)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"100 steps: enqueue {1e4 * (t1 - t0):.1f} us/step, with drain {1e4 * (t2 - t0):.1f} us/step")
for name, kw in (("graph", dict(use_graph=True)), ("stepwise", dict(use_graph=False))):
    runner.run(**kw); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3): runner.run(**kw)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"runner.run() {name}: {1e4 * (t1 - t0) / 3:.1f} us/step incl. reset, store and stats")
    t0 = time.perf_counter()
    for _ in range(3): runner.run(store=False, **kw)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"runner.run(store=False) {name}: {1e4 * (t1 - t0) / 3:.1f} us/step")
runner.reset()
pr = cProfile.Profile(); pr.enable()
for t in range(100): runner.step(t)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(8)
from torch.profiler import profile, ProfilerActivity
runner.reset(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for t in range(100): runner.step(t)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=6, max_name_column_width=50))
