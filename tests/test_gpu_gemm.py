"""The learner's GEMM (include/macjd.h: macjd_gemm; tcgen05 3xTF32 above a size threshold, FP32 SIMT below) against
torch float64 on every operand orientation, ragged sizes, epilogues and split-K."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _gemm(M, N, K, ta, tb, bias=False, act=0, accumulate=False, splitk=False, lda_pad=0, seed=0):
    from macjd_b200 import _native as N_
    L = N_.get_lib()
    g = torch.Generator(device="cuda").manual_seed(seed)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
    A = rn(K, M + lda_pad) if ta else rn(M, K + lda_pad)
    B = rn(N, K + lda_pad) if tb else rn(K, N + lda_pad)
    Am = (A[:, :M].t() if ta else A[:, :K]).double()
    Bm = (B[:, :K].t() if tb else B[:, :N]).double()
    C0 = rn(M, N)
    C = C0.clone()
    b = rn(N) if bias else None
    ws = torch.empty(64 * M * N if splitk else 4, device="cuda") if splitk else None
    L.callv("macjd_gemm", N_.torch_ctx(torch.device("cuda", 0)), M, N, K, A, A.shape[1], int(ta), B, B.shape[1], int(tb), C, N,
            b, act, int(accumulate), ws, ws.numel() if ws is not None else 0)
    ref = Am @ Bm
    if bias:
        ref = ref + b.double()
    if act == 1:
        ref = ref.clamp(min=0)
    elif act == 3:
        ref = torch.sigmoid(ref)
    if accumulate:
        ref = ref + C0.double()
    scale = (Am.abs() @ Bm.abs()).clamp(min=1.0)
    err = ((C.double() - ref).abs() / scale).max().item()
    return err


@pytest.mark.parametrize("ta,tb", [(0, 1), (0, 0), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(1024, 256, 256), (1000, 130, 100), (129, 127, 2051), (4096, 128, 24)])
def test_gemm_orientations_and_ragged_sizes(M, N, K, ta, tb):
    # |error| <= 4e-6 x sum |a||b|: the 3xTF32 split keeps ~2^-20 per product (truncated hi part, lo x lo dropped) and
    # the accumulation over K adds FP32 rounding like any FP32 GEMM (measured 2.5e-6 at K = 2051)
    assert _gemm(M, N, K, ta, tb) < 4e-6
    assert _gemm(M, N, K, ta, tb, lda_pad=3) < 4e-6          # leading dimensions that rule out 16-byte loads


@pytest.mark.parametrize("act", [0, 1, 3])
def test_gemm_epilogues(act):
    assert _gemm(2048, 256, 128, 0, 1, bias=True, act=act) < 4e-6
    assert _gemm(2048, 256, 128, 0, 1, bias=True, act=act, accumulate=True) < 4e-6


def test_gemm_split_k_weight_gradient_shapes():
    # dW = dY^T X: contraction over 20 000 batch rows into one or four tiles
    assert _gemm(128, 128, 20000, 1, 0, splitk=True) < 4e-6
    assert _gemm(256, 262, 20000, 1, 0, splitk=True, accumulate=True) < 4e-6


def test_gemm_small_problems_stay_on_the_fp32_kernel():
    assert _gemm(64, 5, 128, 0, 0, bias=True, act=3) < 1e-6


@pytest.mark.parametrize("tb", [0, 1])
@pytest.mark.parametrize("M,N,K", [(4096, 200, 100), (2048, 768, 256), (1500, 5, 128)])
def test_gemm_presplit_weights_are_bit_identical(M, N, K, tb):
    """Tall products: with workspace the tensor-core kernel fetches B pre-split (csrc/tc_gemm.cuh: tc_pack_b_kernel + one bulk
    copy per stage) instead of splitting it in every CTA -- the same hi / lo values, the same MMAs in the same order."""
    from macjd_b200 import _native as N_
    L = N_.get_lib()
    g = torch.Generator(device="cuda").manual_seed(3)
    A = torch.randn(M, K, device="cuda", generator=g)
    B = torch.randn(N, K, device="cuda", generator=g) if tb else torch.randn(K, N, device="cuda", generator=g)
    b = torch.randn(N, device="cuda", generator=g)
    outs = []
    for ws in (None, torch.empty(((N + 127) // 128) * ((K + 15) // 16) * 4096, device="cuda")):
        C = torch.empty(M, N, device="cuda")
        L.callv("macjd_gemm", N_.torch_ctx(torch.device("cuda", 0)), M, N, K, A, K, 0, B, B.shape[1], tb, C, N, b, 1, 0,
                ws, ws.numel() if ws is not None else 0)
        outs.append(C)
    assert torch.equal(outs[0], outs[1])
    ref = (A.double() @ (B.double().t() if tb else B.double()) + b.double()).clamp(min=0)
    scale = (A.abs().double() @ (B.abs().double().t() if tb else B.abs().double())).clamp(min=1.0)
    assert ((outs[1].double() - ref).abs() / scale).max().item() < 4e-6
