"""Time the learner's GEMM (macjd_gemm) on the shapes of BASELINE config 4 and print TFLOP/s."""
import sys
import torch
sys.path.insert(0, ".")
from macjd_b200 import _native as N

L = N.get_lib()
ctx = N.torch_ctx(torch.device("cuda", 0))
shapes = [("mixer layer  ", 101376, 128, 128, 0, 1), ("qhead fwd    ", 202752, 256, 256, 0, 0), ("gx r|z       ", 204800, 512, 256, 0, 0),
          ("step gh      ", 2048, 512, 256, 0, 0), ("fc1 K=24     ", 204800, 256, 24, 0, 0), ("dW 256x256   ", 256, 256, 202752, 1, 0),
          ("dX           ", 101376, 128, 128, 0, 0)]
only = sys.argv[1] if len(sys.argv) > 1 else None
for name, M, Nn, K, ta, tb in shapes:
    if only and only not in name:
        continue
    A = torch.randn((K, M) if ta else (M, K), device="cuda")
    B = torch.randn((Nn, K) if tb else (K, Nn), device="cuda")
    C = torch.empty(M, Nn, device="cuda")
    ws = torch.empty(64 * M * Nn, device="cuda") if K > 10000 else None
    args = (ctx, M, Nn, K, A, A.shape[1], ta, B, B.shape[1], tb, C, Nn, None, 0, 0, ws, ws.numel() if ws is not None else 0)
    for _ in range(3):
        L.callv("macjd_gemm", *args)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        L.callv("macjd_gemm", *args)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print(f"{name} M={M} N={Nn} K={K} ta={ta} tb={tb}: {us:8.1f} us  {2 * M * Nn * K / us / 1e6:7.1f} TFLOP/s  "
          f"{(M * K + M * Nn + Nn * K) * 4 / us / 1e3:6.0f} GB/s", flush=True)
