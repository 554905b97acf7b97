"""Phase timestamps (SM clock cycles) of one tcgen05 agent CTA; needs a -DMACJD_TC_PROFILE build:
   NVCC_EXTRA=-DMACJD_TC_PROFILE python <pkg>/csrc/build.py --force"""
import ctypes, sys
import torch
sys.path.insert(0, ".")
from tests.agent_checks import random_agent
from macjd_b200 import _native as N
import os
O, A = int(os.environ.get("TC_O", 24)), int(os.environ.get("TC_A", 5))       # TC_O=176 TC_A=33: the C3 shape
mac, _ = random_agent(0, O, A, 128, 128, 2, "cuda")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[3]) if len(sys.argv) > 3 else 3   # T = 1 profiles the cold first step
PATH = int(sys.argv[2]) if len(sys.argv) > 2 else 2
obs = torch.randn(T, M, O, device="cuda")
h = torch.zeros(M, 128, device="cuda")
for _ in range(3):
    mac.agent.run(obs, h, n_steps=T, select=True, test_mode=True, path=PATH)
if len(sys.argv) > 4 and sys.argv[4] == "cold":
    from tools.microbench import flush_l2
    flush_l2()
    mac.agent.run(obs, h, n_steps=T, select=True, test_mode=True, path=PATH)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (64 + 1024))()
N.get_lib().lib.macjd_debug_tc_profile(buf, 64 + 1024)
v = list(buf)
names = {0: "step start", 1: "X written", 2: "D13 ready", 3: "E1 done", 4: "D2 ready", 5: "E2 done", 6: "E3 done", 7: "D4 ready",
         8: "E4 done", 9: "D5 ready", 10: "E5+select done"}
t0 = v[0]
print("epilogue warp 0 (cycles since step start):")
for k in range(11):
    print(f"  {names[k]:>16}: {v[k] - t0:8d}" + (f"  (+{v[k] - v[k-1]})" if k else ""))
inames = {32: "x_full seen", 33: "G1/G3 issued", 34: "a_ready#1", 35: "G2 issued", 36: "a_ready#2", 37: "G4 issued", 38: "a_ready#3", 39: "G5 issued"}
print("issuer:")
for k in range(32, 40):
    print(f"  {inames[k]:>16}: {v[k] - t0:8d}" + (f"  (+{v[k] - v[k-1]})" if k > 32 else ""))


print("CTA 0 kernel entry .. exit (cycles rel. step start):", [v[k] - t0 for k in range(20, 26)])
print("prologue (rel. entry): loads issued", v[26] - v[20], "weights prefetched", v[27] - v[20], "barriers initialised", v[28] - v[20], "constants staged", v[21] - v[20], "cluster barrier", v[22] - v[20], "step start", t0 - v[20])
n_cta = min(512, (M + 63) // 64)
ent = [v[64 + 2 * b] for b in range(n_cta)]
ext = [v[65 + 2 * b] for b in range(n_cta)]
e0 = min(ent)
print(f"{n_cta} CTAs: entry spread {max(ent) - e0} ns; exit min/median/max {min(ext) - e0} / {sorted(ext)[n_cta // 2] - e0} / {max(ext) - e0} ns;"
      f" per-CTA duration min/max {min(x - e for x, e in zip(ext, ent))} / {max(x - e for x, e in zip(ext, ent))} ns")

