"""Agent kernel at the bench's launch shape (availability mask, epsilon-greedy outputs) vs the bare test-mode launch."""
import sys
import torch
sys.path.insert(0, ".")
from tests.agent_checks import random_agent
from tools.microbench import timeit
mac, _ = random_agent(0, 24, 5, 128, 128, 2, "cuda")
M = 8192
obs = torch.randn(1, M, 24, device="cuda")
h = torch.zeros(M, 128, device="cuda")
avail = torch.ones(1, M, 5, dtype=torch.uint8, device="cuda")
hs = torch.zeros(1, M, 128, device="cuda")
out = {"actions": torch.zeros(1, M, dtype=torch.int32, device="cuda"), "power": torch.zeros(1, M, device="cuda")}
cases = {
    "bare test-mode": lambda p: mac.agent.run(obs, h, select=True, test_mode=True, path=p),
    "+ avail": lambda p: mac.agent.run(obs, h, avail=avail, select=True, test_mode=True, path=p),
    "+ avail + out": lambda p: mac.agent.run(obs, h, avail=avail, select=True, test_mode=True, out=out, path=p),
    "+ eps-greedy": lambda p: mac.agent.run(obs, h, avail=avail, select=True, test_mode=False, epsilon=0.3, out=out, path=p),
}
for name, fn in cases.items():
    for p in (3,):
        med, best = timeit(lambda: fn(p), iters=10, warmup=3)
        print(f"{name:>16} path {p}: {med * 1e6:6.1f} us (best {best * 1e6:.1f})", flush=True)
