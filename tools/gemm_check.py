"""Correctness sweep of the tcgen05 GEMM over forced tile configurations (SSCVAE_GEMM_FORCE) against torch.
Usage on the GPU box: python tools/gemm_check.py cfg [cfg ...]"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sscvae  # noqa: E402,F401
from sscvae import _lib  # noqa: E402

SHAPES = [(256, 3600, 1808), (300, 516, 200), (5376, 1000, 64), (256, 4928, 3648), (200, 3600, 4928), (8, 3600, 960), (40, 300, 960), (256, 768, 100), (256, 300, 904), (130, 77, 40), (5376, 1000, 600), (3600, 2048, 5376), (257, 513, 129), (8, 64, 64)]


def main():
    L = _lib.lib()
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    bad = 0
    for cfg in sys.argv[1:]:
        os.environ.pop("SSCVAE_GEMM_SPLITK", None)
        if "/" in cfg:                       # "BN,STAGES/splits"
            cfg, sp = cfg.split("/")
            os.environ["SSCVAE_GEMM_SPLITK"] = sp
        os.environ["SSCVAE_GEMM_FORCE"] = cfg
        for (M, N, K) in SHAPES:
            g = torch.Generator(device="cuda").manual_seed(M + N + K)
            ld = (K + 7) // 8 * 8
            A = torch.zeros(M, ld, device="cuda", dtype=torch.bfloat16)
            B = torch.zeros(N, ld, device="cuda", dtype=torch.bfloat16)
            A[:, :K] = torch.randn(M, K, device="cuda", generator=g).bfloat16()
            B[:, :K] = torch.randn(N, K, device="cuda", generator=g).bfloat16()
            bias = torch.randn(N, device="cuda", generator=g)
            if cfg.startswith("sk"):               # "skS": swapped-operand split-K kernel with S splits (0 = auto)
                if M > 256 or N < 256:
                    continue
                S = int(cfg[2:])
                ldc = (N + 3) // 4 * 4 + 4          # multiple of 4: the TMA-store epilogue is eligible
                Cm = torch.full((M, ldc), -5.0, device="cuda")
                os.environ.pop("SSCVAE_GEMM_FORCE", None)
                _lib.check(L.sscvae_test_gemm_splitk(_lib.ptr(A), ld, _lib.ptr(B), ld, M, N, K, _lib.ptr(Cm), ldc, S,
                                                     _lib.ptr(bias), s))
            else:
                # odd ldc -> staged STG epilogue; ldc % 4 == 0 -> TMA-store epilogue (both must be right)
                ldc = N + 3
                Cm = torch.full((M, ldc), -5.0, device="cuda")
                _lib.check(L.sscvae_test_gemm(_lib.ptr(A), ld, _lib.ptr(B), ld, M, N, K, _lib.ptr(Cm), ldc, _lib.ptr(bias), 0, 0, s))
                ldc2 = (N + 3) // 4 * 4 + 4
                C2 = torch.full((M, ldc2), -5.0, device="cuda")
                _lib.check(L.sscvae_test_gemm(_lib.ptr(A), ld, _lib.ptr(B), ld, M, N, K, _lib.ptr(C2), ldc2, _lib.ptr(bias), 0, 0, s))
                torch.cuda.synchronize()
                if not (torch.equal(C2[:, :N], Cm[:, :N]) and bool((C2[:, N:] == -5.0).all())):
                    bad += 1
                    print(f"[{cfg}] {M}x{N}x{K}: TMA-store epilogue differs from the staged one FAIL", flush=True)
            torch.cuda.synchronize()
            ref = A[:, :K].float() @ B[:, :K].float().t() + bias
            err = (Cm[:, :N] - ref).abs().max().item()
            ok = err <= 2e-3 * (K ** 0.5) and bool((Cm[:, N:] == -5.0).all())
            bad += not ok
            print(f"[{cfg}] {M}x{N}x{K}: max err {err:.3e} {'ok' if ok else 'FAIL'}", flush=True)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
