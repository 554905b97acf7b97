"""Phase timestamps (SM clock cycles) of CTA 0 inside the whole-episode rollout launch (macjd_rollout_steps, C2 shape):
the epilogue warps' and the MMA issuer's stamps of timestep 1 and the env warps' physics window.  Needs a profiling build:
   NVCC_EXTRA=-DMACJD_TC_PROFILE python <pkg>/csrc/build.py --force --out=/root/repo/tools/_prof/libmacjd_prof.so
   MACJD_LIB_PATH=tools/_prof/libmacjd_prof.so python tools/rollout_phase_profile.py [n_envs] [T]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B   # noqa: E402
from macjd_b200 import _native as N   # noqa: E402
from macjd_b200.core.mac import BasicMAC   # noqa: E402
from macjd_b200.runners.episode_runner import BatchedEpisodeRunner   # noqa: E402
from macjd_b200.simulation.environment import ElectromagneticEnvironment   # noqa: E402
from macjd_b200.simulation.scenario import default_spec   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dev = "cuda:0"
rl = B.rl_args(dev, n)
torch.manual_seed(42)
env = ElectromagneticEnvironment(rl, spec=default_spec(n), device=dev, seed=1000)
mac = BasicMAC(B.OBS, rl)
mac.cuda()
runner = BatchedEpisodeRunner(env, mac, None, rl)
for _ in range(2):
    runner.reset()
    runner.rollout(0, T)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
runner.reset()
ev[0].record()
runner.rollout(0, T)
ev[1].record()
torch.cuda.synchronize()
print(f"launch of {T} timesteps, {n} envs: {ev[0].elapsed_time(ev[1]) * 1e3 / T:.2f} us per timestep (warm)")
buf = (ctypes.c_ulonglong * (64 + 1024))()
N.get_lib().lib.macjd_debug_tc_profile(buf, 64 + 1024)
v = list(buf)
names = {0: "step start", 1: "X written", 2: "D13 ready", 3: "E1 done", 4: "D2 ready", 5: "E3 done", 6: "E2 done", 7: "D4 ready",
         8: "E4 done", 9: "D5 ready (views copied)", 11: "selection done", 10: "step end"}
t0 = v[0]
print("epilogue warp 0, timestep 1 (cycles since step start):")
prev = t0
for k in (0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 11, 10):
    print(f"  {names[k]:>24}: {v[k] - t0:8d}  (+{v[k] - prev})")
    prev = v[k]
inames = {32: "x_full seen", 33: "obs products committed", 34: "a_ready#1 (a1) seen", 35: "actor.2 committed", 36: "a_ready#2 (xf) seen",
          37: "input products committed", 38: "a_ready#3 (h') seen", 39: "q.0 committed"}
print("issuer:")
for k in range(32, 40):
    print(f"  {inames[k]:>24}: {v[k] - t0:8d}" + (f"  (+{v[k] - v[k - 1]})" if k > 32 else ""))
print(f"env warps: actions seen {v[40] - t0}, physics done {v[41] - t0}  (+{v[41] - v[40]})")
# CTA lifetimes (%globaltimer ns at entry / exit of every CTA): is the launch as long as its slowest CTA, and how
# far apart are the CTAs?
n_cta = 2 * ((n * 2 + 127) // 128)
ent = [v[64 + 2 * b] for b in range(min(n_cta, 512))]
ext = [v[65 + 2 * b] for b in range(min(n_cta, 512))]
t_first = min(ent)
life = sorted((x - e) / 1e3 for e, x in zip(ent, ext))
print(f"{len(ent)} CTAs: entry spread {(max(ent) - t_first) / 1e3:.1f} us, last exit {(max(ext) - t_first) / 1e3:.1f} us after the first entry")
print(f"CTA lifetime us: min {life[0]:.1f}  median {life[len(life) // 2]:.1f}  max {life[-1]:.1f}   (per timestep: "
      f"{life[0] / T:.2f} / {life[len(life) // 2] / T:.2f} / {life[-1] / T:.2f})")
print("CTA 0 lifetime us:", (ext[0] - ent[0]) / 1e3, " slowest CTAs:", sorted(range(len(ent)), key=lambda b: ent[b] - ext[b])[:8])
