"""Kernel micro-benchmarks on one GPU (CUDA events, warm-up, L2 flush between iterations)."""
import json
import sys
import types

import numpy as np
import torch

sys.path.insert(0, ".")
from macjd_b200.simulation.environment import ElectromagneticEnvironment
from macjd_b200.simulation.scenario import default_spec, scaled_spec
from tests.agent_checks import random_agent

FLUSH = None


def flush_l2():
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    FLUSH.fill_(1)


def timeit(fn, iters=10, warmup=3, flush=True):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return float(np.median(ts)), float(np.min(ts))


def bench_env(out):
    for name, mk, bytes_per in (("default", default_spec, 569.0), ("scaled", lambda n: scaled_spec(n), 8625.0)):
        for n in ((4096, 65536, 1 << 20) if name == "default" else (8192, 65536)):
            for share in (False, True):
                if share and name != "default":
                    continue
                env = ElectromagneticEnvironment(types.SimpleNamespace(), spec=mk(n), device="cuda", share_scenario=share)
                J, R, K = env.num_jammers, env.num_radars, env.num_targets
                act_d = torch.randint(0, 2 * R + 1, (n, J), dtype=torch.int32, device="cuda")
                act_p = torch.rand(n, J, device="cuda")
                noise = torch.rand(n, R * K + J, device="cuda")
                for inj in (True, False):
                    med, best = timeit(lambda: env.step_device(act_d, act_p, noise if inj else None))
                    rec = dict(kernel="env_step", scenario=name, n_envs=n, shared_tables=share, injected_noise=inj,
                               us=med * 1e6, us_best=best * 1e6, env_steps_per_s=n / med,
                               algo_GBps=n * bytes_per / med / 1e9)
                    out.append(rec)
                    print(json.dumps(rec), flush=True)
                del env


def bench_agent(out):
    cfgs = [dict(tag="C2 act", O=24, A=5, H=128, AH=128, Nn=2, M=8192, T=1, flop=279_000),
            dict(tag="C3 act", O=176, A=33, H=128, AH=128, Nn=8, M=65536, T=1, flop=0),
            dict(tag="C1 unroll", O=24, A=5, H=128, AH=128, Nn=2, M=64, T=100, flop=279_000),
            dict(tag="C4 unroll", O=24, A=5, H=256, AH=128, Nn=2, M=2048, T=100, flop=0)]
    for c in cfgs:
        mac, _ = random_agent(0, c["O"], c["A"], c["H"], c["AH"], c["Nn"], "cuda")
        O, A, H, AH, M, T = c["O"], c["A"], c["H"], c["AH"], c["M"], c["T"]
        Op = (O + 31) // 32 * 32
        flop = 2 * (Op * AH + AH * AH + AH * A + Op * H + 6 * H * H + H * H) + A * H * 5
        obs = torch.randn(T, M, O, device="cuda")
        h = torch.zeros(M, H, device="cuda")
        for tile, path in ((32, 1), (64, 1), (0, 2)):
            try:
                fn = lambda: mac.agent.run(obs, h, n_steps=T, select=True, test_mode=True, tile_rows=tile, path=path)
                med, best = timeit(fn, iters=8, warmup=2)
            except Exception as e:  # tile does not fit / path unsupported
                print(c["tag"], tile, path, "skipped:", e)
                continue
            rec = dict(kernel="agent_forward", tag=c["tag"], M=M, T=T, tile=tile, path={1: "simt", 2: "tcgen05"}[path], us=med * 1e6, us_best=best * 1e6,
                       agent_steps_per_s=M * T / med, tflops=M * T * flop / med / 1e12, flop_per_row=flop)
            out.append(rec)
            print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    out = []
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_properties(0).multi_processor_count, "SMs")
    bench_env(out)
    bench_agent(out)
    with open("gpurun_out/microbench.json", "w") as f:
        json.dump(out, f, indent=1)
