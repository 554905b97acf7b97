"""Host-buffer loop (select_actions_host -> step_host) timing; MACJD_DIRECT_HOST_BYTES picks which
buffers the kernels touch in place over PCIe and which go through the copy engines."""
import sys, time, types
import torch
sys.path.insert(0, ".")
from bench import rl_args, OBS, N_AGENTS, N_ACTIONS
from macjd_b200.simulation.environment import ElectromagneticEnvironment
from macjd_b200.simulation.scenario import default_spec
from macjd_b200.core.mac import BasicMAC

n_envs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rl = rl_args("cuda:0", n_envs)
env = ElectromagneticEnvironment(rl, spec=default_spec(n_envs), device="cuda:0", seed=1)
mac = BasicMAC(OBS, rl); mac.cuda()
hb = env.host_buffers()
avail_h = torch.ones(n_envs, N_AGENTS, N_ACTIONS, dtype=torch.uint8).pin_memory()
hb["obs"].copy_(env.get_obs())
mac.init_hidden(n_envs)
def act(t): mac.select_actions_host(hb["obs"], avail_h, t, actions_out=hb["act_d"], power_out=hb["act_p"])
def envs(t): env.step_host(hb)
def both(t): act(t); envs(t)
for name, fn in (("act_host", act), ("env_step_host", envs), ("both", both)):
    for t in range(20): fn(t)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    K = 300
    for t in range(K): fn(t)
    torch.cuda.synchronize()
    print(f"{name:>14}: {(time.perf_counter() - t0) / K * 1e6:7.1f} us/call", flush=True)

# ---- the fused call (BatchedEpisodeRunner.step_host -> macjd_rollout_step_host), with and without the
# observation write-back and with / without env groups (MACJD_HOST_GROUP_ENVS is read per call)
import os
from macjd_b200.runners.episode_runner import BatchedEpisodeRunner
runner = types.SimpleNamespace(mac=mac, env=env, t_env=0)
hb_noobs = {k: v for k, v in hb.items() if k != "obs"}
obs_in = hb["obs"].clone().pin_memory()
for groups in ("0", "2048", "1024"):
    os.environ["MACJD_HOST_GROUP_ENVS"] = groups
    for name, host in (("fused", hb), ("fused, no obs out", hb_noobs)):
        fn = lambda t: BatchedEpisodeRunner.step_host(runner, obs_in, avail_h, host)
        for t in range(20): fn(t)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in range(K): fn(t)
        torch.cuda.synchronize()
        print(f"groups {groups:>5} {name:>18}: {(time.perf_counter() - t0) / K * 1e6:7.1f} us/call", flush=True)
