"""tcgen05 building blocks (csrc/tc05.cuh): 3xTF32 GEMM self-test against float64."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(64, 128, 32), (64, 128, 128), (64, 256, 64), (128, 128, 64), (128, 256, 32), (64, 16, 8)])
def test_3xtf32_gemm_matches_float64(M, N, K):
    from macjd_b200 import _native as Nat
    lib = Nat.get_lib()
    g = torch.Generator(device="cuda").manual_seed(M * 1000 + N + K)
    A = torch.randn(M, K, device="cuda", generator=g) * 3
    B = torch.randn(N, K, device="cuda", generator=g)
    D = torch.full((M, N), float("nan"), device="cuda")
    lib.callv("macjd_tc_gemm_selftest", Nat.torch_ctx("cuda:0"), M, N, K, A, B, D)
    torch.cuda.synchronize()
    ref = (A.double() @ B.double().t())
    scale = (A.double().abs() @ B.double().abs().t())            # magnitude of the summed terms
    err = ((D.double() - ref).abs() / scale).max().item()
    assert err < 2e-6, f"relative-to-magnitude error {err:.3e} (plain TF32 would be ~1e-3)"


@pytest.mark.parametrize("N,K", [(128, 32), (64, 64)])
def test_fragment_layout_tmem_load(N, K):
    """tcgen05.ld.16x256b on an M = 64 accumulator: same result as the lane-per-thread read."""
    from macjd_b200 import _native as Nat
    lib = Nat.get_lib()
    g = torch.Generator(device="cuda").manual_seed(N + K)
    A = torch.randn(64, K, device="cuda", generator=g)
    B = torch.randn(N, K, device="cuda", generator=g)
    D0 = torch.full((64, N), float("nan"), device="cuda")
    D1 = torch.full((64, N), float("nan"), device="cuda")
    lib.callv("macjd_tc_gemm_selftest", Nat.torch_ctx("cuda:0"), 64, N, K, A, B, D0)
    lib.callv("macjd_tc_gemm_selftest", Nat.torch_ctx("cuda:0"), 64, N | (1 << 30), K, A, B, D1)
    torch.cuda.synchronize()
    assert torch.equal(D0, D1)
