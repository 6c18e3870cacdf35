"""K sweep of the tcgen05 GEMM: separates the fixed per-launch cost from the per-k-block cost.
Back-to-back launches (no sync in between), operands L2-resident."""
import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sscvae
from sscvae import _lib
L = _lib.lib()
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)

def run(M, N, K, cfg, iters=200):
    if cfg: os.environ["SSCVAE_GEMM_FORCE"] = cfg
    else: os.environ.pop("SSCVAE_GEMM_FORCE", None)
    A = torch.randn(M, K, device="cuda").bfloat16(); B = torch.randn(N, K, device="cuda").bfloat16()
    Cm = torch.zeros(M, N, device="cuda")
    fn = lambda: _lib.check(L.sscvae_test_gemm(_lib.ptr(A), K, _lib.ptr(B), K, M, N, K, _lib.ptr(Cm), N, None, 0, 0, s))
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / iters * 1e3
    f = lambda: torch.matmul(A, B.t())
    for _ in range(5): f()
    e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize()
    return t, e0.elapsed_time(e1) / iters * 1e3

for (M, N) in [(128, 64), (256, 3600), (256, 900)]:
    for cfg in ["64,4", "64,8", "128,3"]:
        out = []
        for K in [64, 256, 1024, 2048, 4096, 8192]:
            t, tb = run(M, N, K, cfg)
            out.append(f"K={K}: {t:6.1f} (cublas {tb:5.1f})")
        print(f"{M}x{N} [{cfg}] " + " | ".join(out), flush=True)
