"""env_step_kernel at the bench's batch size (and one large size) under the MACJD_ENV_BLOCK / MACJD_ENV_STAGE knobs."""
import sys, types
import torch
sys.path.insert(0, ".")
from tools.microbench import timeit
from macjd_b200.simulation.environment import ElectromagneticEnvironment
from macjd_b200.simulation.scenario import default_spec
for n in (4096, 65536, 1 << 20):
    env = ElectromagneticEnvironment(types.SimpleNamespace(), spec=default_spec(n), device="cuda")
    act_d = torch.randint(0, 5, (n, 2), dtype=torch.int32, device="cuda")
    act_p = torch.rand(n, 2, device="cuda")
    med, best = timeit(lambda: env.step_device(act_d, act_p), iters=10, warmup=3)
    medw, bestw = timeit(lambda: env.step_device(act_d, act_p), iters=10, warmup=3, flush=False)
    print(f"n={n}: {med * 1e6:8.1f} us (best {best * 1e6:.1f}) after an L2 flush; {medw * 1e6:8.1f} us (best {bestw * 1e6:.1f}) warm", flush=True)
