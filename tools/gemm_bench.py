"""Micro-benchmark of the tcgen05 GEMM at the shapes of the training step against cuBLAS (torch.matmul),
over tile configurations (SSCVAE_GEMM_FORCE). Usage on the GPU box: python tools/gemm_bench.py"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sscvae  # noqa: E402
from sscvae import _lib  # noqa: E402

SHAPES = [  # (M, N, K, what)
    (256, 3600, 1808, "att gates"), (256, 3600, 4760, "enc gates"), (256, 768, 904, "q"), (256, 300, 904, "fc"),
    (256, 4008, 3600, "bwd dXEZ"), (256, 4760, 3600, "bwd dXEH"), (256, 1808, 3600, "bwd dXA"), (256, 900, 768, "bwd dh1_q"),
    (3600, 2048, 5376, "wgrad"), (5376, 10000, 600, "vocab"), (5376, 3600, 600, "emb gates"),
]
CONFIGS = [None, "256,4", "128,6", "128,3", "64,8", "64,4", "32,10", "32,5", "16,6"]


def timeit(fn, iters=30, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3


def main():
    L = _lib.lib()
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # 256 MB > L2
    for (M, N, K, what) in SHAPES:
        A = torch.randn(M, K, device="cuda").bfloat16()
        B = torch.randn(N, K, device="cuda").bfloat16()
        Cm = torch.zeros(M, N, device="cuda")
        ref_hot = timeit(lambda: torch.matmul(A, B.t()))
        ref_cold = timeit(lambda: torch.matmul(A, B.t()), flush=flush)
        line = f"{what:10s} {M}x{N}x{K}: cublas hot {ref_hot:7.1f} cold {ref_cold:7.1f} us |"
        for cfg in CONFIGS:
            if cfg is None:
                os.environ.pop("SSCVAE_GEMM_FORCE", None)
            else:
                os.environ["SSCVAE_GEMM_FORCE"] = cfg
            fn = lambda: _lib.check(L.sscvae_test_gemm(_lib.ptr(A), K, _lib.ptr(B), K, M, N, K, _lib.ptr(Cm), N, None, 0, 0, s))
            try:
                hot = timeit(fn)
                cold = timeit(fn, flush=flush)
                line += f" [{cfg or 'auto'}] {hot:6.1f}/{cold:6.1f}"
            except Exception as e:
                line += f" [{cfg}] ERR"
        print(line, flush=True)
    os.environ.pop("SSCVAE_GEMM_FORCE", None)


if __name__ == "__main__":
    main()
