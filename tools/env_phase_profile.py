"""%globaltimer stamps of env block 0 / thread 0 inside a rollout step (needs a -DMACJD_TC_PROFILE build)."""
import ctypes, sys, torch
sys.path.insert(0, ".")
import bench
from tools.microbench import flush_l2
from macjd_b200 import _native as N
from macjd_b200.simulation.environment import ElectromagneticEnvironment
from macjd_b200.simulation.scenario import default_spec
from macjd_b200.core.mac import BasicMAC
from macjd_b200.utils.replay_buffer import EpisodeReplayBuffer
from macjd_b200.runners.episode_runner import BatchedEpisodeRunner
dev = "cuda:0"; n_envs = 4096
rl = bench.rl_args(dev, n_envs)
env = ElectromagneticEnvironment(rl, spec=default_spec(n_envs), device=dev, seed=1)
mac = BasicMAC(bench.OBS, rl); mac.cuda()
runner = BatchedEpisodeRunner(env, mac, EpisodeReplayBuffer(rl, device=dev), rl)
runner.reset()
lib = N.get_lib().lib
names = ["at the wait", "past the wait", "step counter read", "jammer loop done", "radar loop done", "workers joined", "outputs written"]
for label, fl in (("warm", False), ("after an L2 flush", True)):
    for t in range(5):
        if fl: flush_l2()
        runner.step(t)
    torch.cuda.synchronize()
    a = (ctypes.c_ulonglong * (64 + 1024))(); e = (ctypes.c_ulonglong * 16)()
    lib.macjd_debug_tc_profile(a, 64 + 1024); lib.macjd_debug_env_profile(e)
    ent = [a[64 + 2 * b] for b in range(128)]; ext = [a[65 + 2 * b] for b in range(128)]
    t0 = min(ent)
    print(f"{label}: agent last CTA exit {max(ext) - t0} ns; env block 0: " + ", ".join(f"{n} {e[i] - t0}" for i, n in enumerate(names)))
