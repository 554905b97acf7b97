"""Per-launch time of the two rollout kernels back to back (warm L2 / instruction caches) vs after an L2 flush."""
import sys
import torch
sys.path.insert(0, ".")
import bench
from tools.microbench import flush_l2
from macjd_b200.simulation.environment import ElectromagneticEnvironment
from macjd_b200.simulation.scenario import default_spec
from macjd_b200.core.mac import BasicMAC
from macjd_b200.utils.replay_buffer import EpisodeReplayBuffer
from macjd_b200.runners.episode_runner import BatchedEpisodeRunner
dev = "cuda:0"
n_envs = 4096
rl = bench.rl_args(dev, n_envs)
env = ElectromagneticEnvironment(rl, spec=default_spec(n_envs), device=dev, seed=1)
mac = BasicMAC(bench.OBS, rl); mac.cuda()
buf = EpisodeReplayBuffer(rl, device=dev)
runner = BatchedEpisodeRunner(env, mac, buf, rl)
runner.reset(); runner.step(0)
lib, ctx = mac.agent.lib(), mac.agent._ctx()
aio, eio, wts = runner._agent_io[0], runner._env_io[0], mac.agent.packed().cstruct()
calls = {"agent": lambda: lib.call("macjd_agent_forward", ctx, wts, aio), "env": lambda: lib.call("macjd_env_step", ctx, env._ctab, eio)}
for name, fn in calls.items():
    for _ in range(5): fn()
    torch.cuda.synchronize()
    N = 200
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush_l2()                      # keeps the GPU busy while the launches queue up
    a.record()
    for _ in range(N): fn()
    b.record(); torch.cuda.synchronize()
    warm = a.elapsed_time(b) * 1e3 / N
    cold = []
    for _ in range(10):
        flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        cold.append(a.elapsed_time(b) * 1e3)
    print(f"{name}: {warm:.1f} us per launch back to back (warm), {sorted(cold)[5]:.1f} us after an L2 flush", flush=True)
def both():
    calls["agent"](); calls["env"]()
N = 200
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
flush_l2(); a.record()
for _ in range(N): both()
b.record(); torch.cuda.synchronize()
print(f"agent + env: {a.elapsed_time(b) * 1e3 / N:.1f} us per step back to back (warm)")
