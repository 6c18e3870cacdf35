"""GPU: the tcgen05 / TMA GEMM against torch on the same bf16 operands."""
import ctypes as C

import pytest
import torch

from sscvae import _lib

pytestmark = pytest.mark.gpu

SHAPES = [
    (128, 64, 64), (256, 3600, 1808), (256, 768, 904), (256, 300, 904), (5376, 3600, 600), (9216, 768, 2048),
    (37, 50, 72), (1, 8, 8), (300, 129, 1000), (3600, 1808, 5376), (256, 3856, 3600), (130, 10000, 600),
]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_matches_torch(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    lda, ldb = (K + 7) // 8 * 8 + 8, (K + 7) // 8 * 8
    A = torch.zeros(M, lda, device="cuda", dtype=torch.bfloat16)
    B = torch.zeros(N, ldb, device="cuda", dtype=torch.bfloat16)
    A[:, :K] = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    B[:, :K] = torch.randn(N, K, device="cuda", generator=g).bfloat16()
    A[:, K:] = 7.0            # garbage beyond K must not be read
    bias = torch.randn(N, device="cuda", generator=g)
    ldc = N + 3
    Cm = torch.full((M, ldc), -5.0, device="cuda")
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(_lib.lib().sscvae_test_gemm(_lib.ptr(A), lda, _lib.ptr(B), ldb, M, N, K, _lib.ptr(Cm), ldc, _lib.ptr(bias), 0, 0, s))
    torch.cuda.synchronize()
    ref = A[:, :K].float() @ B[:, :K].float().t() + bias
    err = (Cm[:, :N] - ref).abs().max().item()
    assert err <= 2e-3 * (K ** 0.5), (err, M, N, K)
    assert (Cm[:, N:] == -5.0).all()          # nothing written outside the tile's valid columns
    # accumulate + tanh epilogue
    _lib.check(_lib.lib().sscvae_test_gemm(_lib.ptr(A), lda, _lib.ptr(B), ldb, M, N, K, _lib.ptr(Cm), ldc, None, 0, 1, s))
    torch.cuda.synchronize()
    ref2 = ref + A[:, :K].float() @ B[:, :K].float().t()
    assert (Cm[:, :N] - ref2).abs().max().item() <= 4e-3 * (K ** 0.5)
    _lib.check(_lib.lib().sscvae_test_gemm(_lib.ptr(A), lda, _lib.ptr(B), ldb, M, N, K, _lib.ptr(Cm), ldc, None, 1, 0, s))
    torch.cuda.synchronize()
    ref3 = torch.tanh(A[:, :K].float() @ B[:, :K].float().t())
    assert (Cm[:, :N] - ref3).abs().max().item() <= 2e-2
