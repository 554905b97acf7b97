"""Randomised repeat-and-compare stress of the GPU kernels (compute-sanitizer is not available on the pool):
every launch is issued twice on identical inputs and must reproduce bit for bit (a shared-memory or TMEM
race shows up as a mismatch), and the tensor-core paths must stay within their stated bound of the FP32 kernel."""
import sys, types
import numpy as np
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from tests import agent_checks as AC
from macjd_b200.simulation.environment import ElectromagneticEnvironment
from macjd_b200.simulation.scenario import hetero_spec, scaled_spec



def run(seed=0, n_iter=40):
  rng = np.random.default_rng(seed)
  bad = 0
  for it in range(n_iter):
      A = int(rng.integers(1, 9)); O = int(rng.choice([8, 24, 30, 40, 100]))
      M = int(rng.choice([1, 63, 64, 65, 128, 129, 500, 4096, 8192, 9000])); T = int(rng.choice([1, 1, 2, 4]))
      mac, _ = AC.random_agent(int(rng.integers(1 << 30)), O, A, 128, 128, 2, "cuda")
      g = torch.Generator(device="cuda").manual_seed(int(rng.integers(1 << 30)))
      obs = torch.randn(T, M, O, device="cuda", generator=g) * float(rng.choice([0.5, 3.0, 30.0]))
      h0 = torch.randn(M, 128, device="cuda", generator=g) * 0.5
      avail = torch.rand(T, M, A, device="cuda", generator=g) < 0.7
      avail[..., 0] = True
      kw = dict(n_steps=T, avail=avail, select=True, test_mode=bool(rng.integers(2)), epsilon=0.3, seed=7, rng_step=11,
                want_q=True, want_params=True, want_greedy=True, want_hidden_seq=True)
      outs = {}
      for name, extra in (("simt", dict(path=1)), ("pair", dict(path=3, split_unroll=False)), ("pair2", dict(path=3, split_unroll=False)),
                          ("split", dict(path=3)), ("split2", dict(path=3))):
          if name.startswith("split") and not kw["test_mode"]:
              continue
          outs[name] = mac.agent.run(obs, h0.clone(), **kw, **extra)
      torch.cuda.synchronize()
      for a, b in (("pair", "pair2"), ("split", "split2")):
          if a in outs:
              for k in outs[a]:
                  if not torch.equal(outs[a][k], outs[b][k]):
                      bad += 1; print(f"iter {it}: {a} not reproducible in {k} (M={M} T={T} A={A} O={O})", flush=True)
      scale = max(1.0, float(outs["simt"]["q_all"].abs().max()))
      for name in ("pair", "split"):
          if name in outs:
              for k, tol in (("q_all", 3e-4), ("hidden_seq", 1e-4), ("params_all", 1e-5)):
                  err = float((outs[name][k] - outs["simt"][k]).abs().max())
                  if not err <= tol * scale:
                      bad += 1; print(f"iter {it}: {name} {k} off by {err} (scale {scale}; M={M} T={T} A={A} O={O})", flush=True)
  for it in range(max(4, n_iter // 5)):
      n = int(rng.choice([1, 100, 129, 4096, 16000, 20000])); R = int(rng.choice([1, 2, 3])); J = int(rng.choice([1, 2, 3])); K = int(rng.choice([1, 2]))
      spec = scaled_spec(n, n_jammers=J, n_radars=R, n_targets=K, seed=int(rng.integers(1 << 30)), episode_limit=7)
      outs = []
      for rep in range(2):
          env = ElectromagneticEnvironment(types.SimpleNamespace(), spec=spec, device="cuda", seed=5)
          g = torch.Generator(device="cuda").manual_seed(3)
          act_d = torch.randint(0, 2 * R + 1, (n, J), dtype=torch.int32, device="cuda", generator=g)
          act_p = torch.rand(n, J, device="cuda", generator=g)
          env.step_device(act_d, act_p); env.step_device(act_d, act_p)
          torch.cuda.synchronize()
          outs.append({k: getattr(env, k).clone() for k in ("reward64", "pd", "detected", "tracking", "obs", "state", "terminated", "step_count", "r_j")})
      for k in outs[0]:
          if not torch.equal(outs[0][k], outs[1][k]):
              bad += 1; print(f"env iter {it}: {k} not reproducible (n={n} J={J} R={R} K={K})", flush=True)
  print("stress done:", "OK" if bad == 0 else f"{bad} problems")
  return bad


if __name__ == "__main__":
    sys.exit(1 if run(int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 40) else 0)
