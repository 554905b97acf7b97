"""Where the env step's time goes at the bench size: views only (reset), physics only, both."""
import sys, types
import torch
sys.path.insert(0, ".")
from tools.microbench import timeit
from macjd_b200.simulation.environment import ElectromagneticEnvironment
from macjd_b200.simulation.scenario import default_spec
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = ElectromagneticEnvironment(types.SimpleNamespace(), spec=default_spec(n), device="cuda")
act_d = torch.randint(0, 5, (n, 2), dtype=torch.int32, device="cuda")
act_p = torch.rand(n, 2, device="cuda")
noise = torch.rand(n, 4, device="cuda")
extras = ("r_d", "r_p", "r_j", "reward64", "pd", "detected", "tracking", "snr0", "snr1", "jsr_db", "pd_net", "jam_power")
cases = {
    "step (all outputs)": lambda: env.step_device(act_d, act_p),
    "step, injected noise": lambda: env.step_device(act_d, act_p, noise),
    "step, no views": lambda: env.step_device(act_d, act_p, out={"state": None, "obs": None, "avail": None}),
    "step, no extras": lambda: env.step_device(act_d, act_p, out={k: None for k in extras}),
    "step, reward only": lambda: env.step_device(act_d, act_p, out={k: None for k in extras + ("state", "obs", "avail")}),
    "reset (views only)": lambda: env._reset_device(),
}
for name, fn in cases.items():
    med, best = timeit(fn, iters=10, warmup=3)
    print(f"n={n} {name:>22}: {med * 1e6:7.1f} us (best {best * 1e6:.1f})", flush=True)
