"""Tiny driver for ncu captures: runs one kernel configuration a few times."""
import sys
import types

import torch

sys.path.insert(0, ".")
which = sys.argv[1] if len(sys.argv) > 1 else "agent"
if which == "agent":
    from tests.agent_checks import random_agent
    mac, _ = random_agent(0, 24, 5, 128, 128, 2, "cuda")
    M = 8192
    obs = torch.randn(1, M, 24, device="cuda")
    h = torch.zeros(M, 128, device="cuda")
    for _ in range(4):
        if len(sys.argv) > 2 and sys.argv[2] in ("tc", "tc2"):
            mac.agent.run(obs, h, select=True, test_mode=True, path=3)
        else:
            mac.agent.run(obs, h, select=True, test_mode=True, tile_rows=int(sys.argv[2]) if len(sys.argv) > 2 else 0, path=1)
elif which == "env":
    from macjd_b200.simulation.environment import ElectromagneticEnvironment
    from macjd_b200.simulation.scenario import default_spec
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
    env = ElectromagneticEnvironment(types.SimpleNamespace(), spec=default_spec(n), device="cuda")
    act_d = torch.randint(0, 5, (n, 2), dtype=torch.int32, device="cuda")
    act_p = torch.rand(n, 2, device="cuda")
    noise = torch.rand(n, 4, device="cuda")
    for _ in range(4):
        env.step_device(act_d, act_p, noise)
torch.cuda.synchronize()
print("done")
