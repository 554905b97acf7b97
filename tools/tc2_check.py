"""Check the 2-CTA (cta_group::2) agent kernel (path 3) against the SIMT kernel and time it."""
import json, sys
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from tests import agent_checks as AC
from tools.microbench import timeit

res = {}
for M, T in ((128, 1), (100, 3), (8192, 1), (8200, 2)):
    mac, args = AC.random_agent(7, 24, 5, 128, 128, 2, "cuda")
    g = torch.Generator(device="cuda").manual_seed(3)
    obs = torch.randn(T, M, 24, device="cuda", generator=g) * 5
    h0 = torch.randn(M, 128, device="cuda", generator=g) * 0.5
    avail = torch.rand(T, M, 5, device="cuda", generator=g) < 0.7
    avail[..., 0] = True
    out = {}
    for path in (1, 2, 3):
        h = h0.clone()
        out[path] = mac.agent.run(obs, h, n_steps=T, avail=avail, select=True, test_mode=True, want_q=True,
                                  want_params=True, want_greedy=True, want_hidden_seq=True, path=path)
        torch.cuda.synchronize()
    a = out[1]
    for path in (2, 3):
        b = out[path]
        errs = {k: float((b[k] - a[k]).abs().max()) for k in ("params_all", "hidden_seq", "q_all", "hidden")}
        same = float((a["actions"] == b["actions"]).float().mean())
        print(M, T, "path", path, errs, "actions equal", same, flush=True)
        res[f"{M}x{T}_p{path}"] = dict(errs=errs, actions_equal=same)

mac, _ = AC.random_agent(0, 24, 5, 128, 128, 2, "cuda")
for M in (8192, 18944, 65536):
    obs = torch.randn(1, M, 24, device="cuda")
    h = torch.zeros(M, 128, device="cuda")
    for path in (1, 2, 3):
        fn = lambda: mac.agent.run(obs, h, n_steps=1, select=True, test_mode=True, path=path)
        med, best = timeit(fn, iters=10, warmup=3)
        print("M", M, "path", path, "us", med * 1e6, "best", best * 1e6, flush=True)
        res[f"time_M{M}_p{path}"] = dict(us=med * 1e6, best=best * 1e6)
json.dump(res, open("gpurun_out/tc2_check.json", "w"), indent=1)
