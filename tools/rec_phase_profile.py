"""Phase timestamps (SM clock cycles) of CTA 0 inside the one-launch GRU recurrence (csrc/gru_rec_tc2.cuh) at the C4 shape
(M = 2048 rows, H = 256, T = 100).  Needs the profiling build (see tools/rollout_phase_profile.py):
   MACJD_LIB_PATH=tools/_prof/libmacjd_prof.so python tools/rec_phase_profile.py [M] [T]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from macjd_b200 import _native as N   # noqa: E402
from tests.agent_checks import random_agent   # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
mac, _ = random_agent(0, 24, 5, 256, 128, 2, "cuda")
obs = torch.randn(T, M, 24, device="cuda")
for _ in range(2):
    mac.agent.run(obs, None, n_steps=T, zero_init=True, want_hidden_seq=True, path=0)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (64 + 1024))()
N.get_lib().lib.macjd_debug_tc_profile(buf, 64 + 1024)
v = list(buf)
t0 = v[0]
names = {0: "step start", 1: "block 0 ready", 2: "block 0 gates done", 3: "block 1 ready", 4: "block 1 gates done", 5: "h' written, arrived"}
print("epilogue warp 0, timestep 1 (cycles since step start):")
prev = t0
for k in range(6):
    print(f"  {names[k]:>24}: {v[k] - t0:8d}  (+{v[k] - prev})")
    prev = v[k]
print("issuer:")
for k, nm in ((32, "a_ready seen"), (36, "first stage issued"), (33, "block 0 issued"), (34, "block 1 issued")):
    print(f"  {nm:>24}: {v[k] - t0:8d}")
