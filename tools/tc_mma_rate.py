import ctypes, sys
import torch
sys.path.insert(0, ".")
from macjd_b200 import _native as N
lib = N.get_lib()
out = torch.zeros(2, dtype=torch.int64, device="cuda")
for M, Nn in ((64, 64), (64, 128), (64, 256), (128, 64), (128, 128), (128, 256)):
    for n in (1, 16, 64):
        for _ in range(2):
            lib.lib.macjd_debug_tc_mma_rate(ctypes.byref(N.torch_ctx("cuda:0")), M, Nn, n, ctypes.c_void_p(out.data_ptr()))
            torch.cuda.synchronize()
        issue, done = out.tolist()
        print(f"M={M:3d} N={Nn:3d} n={n:3d}: issue {issue:6d} cyc ({issue / n:6.1f}/mma)  complete {done:6d} cyc ({done / n:6.1f}/mma)  MAC/clk {M * Nn * 8 * n / done:7.1f}")

print("cta_group::2 (pair M):")
for M, Nn in ((128, 64), (128, 128), (128, 256), (256, 128), (256, 256)):
    for n in (1, 16, 64):
        for _ in range(2):
            lib.lib.macjd_debug_tc2_mma_rate(ctypes.byref(N.torch_ctx("cuda:0")), M, Nn, n, ctypes.c_void_p(out.data_ptr()))
            torch.cuda.synchronize()
        issue, done = out.tolist()
        print(f"M={M:3d} N={Nn:3d} n={n:3d}: issue {issue:6d} cyc ({issue / n:6.1f}/mma)  complete {done:6d} cyc ({done / n:6.1f}/mma)  MAC/clk/SM {M * Nn * 8 * n / done / 2:7.1f}")

print("several issuing warps (cta_group::1), separate accumulators:")
for M, Nn in ((64, 64), (128, 64), (64, 128), (128, 128)):
    for issuers in (1, 2, 3, 4):
        n = 64
        for _ in range(2):
            lib.lib.macjd_debug_tc_mma_rate_multi(ctypes.byref(N.torch_ctx("cuda:0")), M, Nn, n, issuers, ctypes.c_void_p(out.data_ptr()))
            torch.cuda.synchronize()
        done = out.tolist()[0]
        print(f"M={M:3d} N={Nn:3d} issuers={issuers} n={n}: complete {done:6d} cyc ({done / (n * issuers):6.1f}/mma)")
