"""Event-timed duration of the tcgen05 agent kernels next to the CTA lifetimes the kernel records itself
(-DMACJD_TC_PROFILE build): separates launch ramp / teardown from work."""
import ctypes, sys
import torch
sys.path.insert(0, ".")
from tests.agent_checks import random_agent
from tools.microbench import flush_l2
from macjd_b200 import _native as N
mac, _ = random_agent(0, 24, 5, 128, 128, 2, "cuda")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
obs = torch.randn(1, M, 24, device="cuda")
h = torch.zeros(M, 128, device="cuda")
big = torch.ones(256 << 20, dtype=torch.uint8, device="cuda")
for path in (2, 3):
    for mode in ("flush", "noflush", "flush+read", "flush+wait"):
        ev = []
        for it in range(6):
            if mode != "noflush":
                flush_l2()
            else:
                torch.cuda.synchronize()
            if mode == "flush+read":
                sink = big.sum()           # evicts the dirty lines the flush left behind: L2 cold AND clean
            if mode == "flush+wait":
                torch.cuda._sleep(400000)  # ~200 us of idle SMs after the flush
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            mac.agent.run(obs, h, n_steps=1, select=True, test_mode=True, path=path)
            b.record()
            torch.cuda.synchronize()
            ev.append(a.elapsed_time(b) * 1e3)
        buf = (ctypes.c_ulonglong * (64 + 1024))()
        N.get_lib().lib.macjd_debug_tc_profile(buf, 64 + 1024)
        v = list(buf)
        n_cta = min(512, (M + 63) // 64)
        ent = [v[64 + 2 * b] for b in range(n_cta)]
        ext = [v[65 + 2 * b] for b in range(n_cta)]
        print(f"path {path} {mode:>10}: events {sorted(ev)[len(ev) // 2]:.1f} us (min {min(ev):.1f}); first entry -> last exit {(max(ext) - min(ent)) / 1e3:.1f} us", flush=True)
