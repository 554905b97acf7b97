"""Single-CTA (path 2) vs CTA-pair (path 3) tcgen05 agent kernel on the learner's unroll shape."""
import sys
import torch
sys.path.insert(0, ".")
from tests.agent_checks import random_agent
from tools.microbench import timeit
mac, _ = random_agent(0, 24, 5, 128, 128, 2, "cuda")
for M, T in ((64, 100), (32, 100), (128, 100), (64, 1)):
    obs = torch.randn(T, M, 24, device="cuda")
    h = torch.zeros(M, 128, device="cuda")
    for path in (2, 3):
        fn = lambda: mac.agent.run(obs, h, n_steps=T, zero_init=True, want_q=True, want_greedy=True, want_hidden_seq=True, path=path)
        med, best = timeit(fn, iters=5, warmup=2)
        print(f"M={M} T={T} path {path}: {med * 1e6:9.1f} us  ({med * 1e6 / T:6.1f} us/step)", flush=True)
