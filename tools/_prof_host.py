import sys, cProfile, pstats, types, torch
sys.path.insert(0, ".")
from bench import rl_args, OBS, N_AGENTS, N_ACTIONS
from macjd_b200.simulation.environment import ElectromagneticEnvironment
from macjd_b200.simulation.scenario import default_spec
from macjd_b200.core.mac import BasicMAC
n_envs = 64
rl = rl_args("cuda:0", n_envs)
env = ElectromagneticEnvironment(rl, spec=default_spec(n_envs), device="cuda:0", seed=1)
mac = BasicMAC(OBS, rl); mac.cuda()
hb = env.host_buffers()
avail_h = torch.ones(n_envs, N_AGENTS, N_ACTIONS, dtype=torch.uint8).pin_memory()
hb["obs"].copy_(env.get_obs())
mac.init_hidden(n_envs)
def act(t): mac.select_actions_host(hb["obs"], avail_h, t, actions_out=hb["act_d"], power_out=hb["act_p"])
for t in range(50): act(t)
pr = cProfile.Profile(); pr.enable()
for t in range(2000): act(t)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
