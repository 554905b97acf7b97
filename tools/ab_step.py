"""A/B helper: per-timestep time of the whole-episode rollout launch, of the single-step fused launch (L2 flushed) and of the
host-buffer step, for the library named by MACJD_LIB_PATH.   python tools/ab_step.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B   # noqa: E402
from macjd_b200.core.mac import BasicMAC   # noqa: E402
from macjd_b200.runners.episode_runner import BatchedEpisodeRunner   # noqa: E402
from macjd_b200.simulation.environment import ElectromagneticEnvironment   # noqa: E402
from macjd_b200.simulation.scenario import default_spec   # noqa: E402

n, T, dev = 4096, 100, "cuda:0"
rl = B.rl_args(dev, n)
torch.manual_seed(42)
env = ElectromagneticEnvironment(rl, spec=default_spec(n), device=dev, seed=1000)
mac = BasicMAC(B.OBS, rl)
mac.cuda()
runner = BatchedEpisodeRunner(env, mac, None, rl)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(2):
    runner.reset(); runner.rollout(0, T)
res = []
for rep in range(5):
    runner.reset()
    flush.fill_(rep)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); runner.rollout(0, T); b.record(); torch.cuda.synchronize()
    res.append(a.elapsed_time(b) * 1e3 / T)
print("episode launch, us per timestep:", " ".join(f"{x:.2f}" for x in res))
runner.reset()
for t in range(5):
    runner.step(t)
ts = []
for t in range(5, 65):
    flush.fill_(t & 255)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); runner.step(t); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
ts.sort()
print(f"single-step fused launch, L2 flushed: median {ts[len(ts) // 2]:.2f} us  min {ts[0]:.2f}")
ts = []
runner.reset()
for t in range(5):
    runner.step(t)
torch.cuda.synchronize()
for t in range(5, 65):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); runner.step(t); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
ts.sort()
print(f"single-step fused launch, warm: median {ts[len(ts) // 2]:.2f} us  min {ts[0]:.2f}")
hb = env.host_buffers()
hb_s = {k: v for k, v in hb.items() if k != "obs"}
hb_s["state"] = torch.zeros(n, env.state_dim, dtype=torch.float32).pin_memory()
avail_h = torch.ones(n, env.num_jammers, mac.agent.n_actions, dtype=torch.uint8).pin_memory()
env.reset(); hb_s["state"].copy_(env.get_state()); mac.init_hidden(n)
runner.t_env = 0
for _ in range(20):
    runner.step_host(hb_s["state"], avail_h, hb_s)
torch.cuda.synchronize()
hs = []
for rep in range(8):
    t0 = time.perf_counter()
    for _ in range(200):
        runner.step_host(hb_s["state"], avail_h, hb_s)
    torch.cuda.synchronize()
    hs.append((time.perf_counter() - t0) / 200 * 1e6)
hs.sort()
print(f"host step (state rows), 8 x 200 steps: min {hs[0]:.1f}  median {hs[4]:.1f}  max {hs[-1]:.1f} us")
