import ctypes as C, os, sys, torch
sys.path.insert(0, "/root/repo")
import sscvae
from sscvae import _lib
L=_lib.lib(); s=C.c_void_p(torch.cuda.current_stream().cuda_stream)
os.environ["SSCVAE_GEMM_DBG"]=sys.argv[1]
for (M,N,K) in ((256,3600,1920),(256,768,960)):
    A=torch.randn(M,K,device="cuda").bfloat16(); B=torch.randn(N,K,device="cuda").bfloat16()
    Cm=torch.zeros(M,N,device="cuda")
    for S in (1,4):
        print("shape",M,N,K,"S",S, flush=True)
        for it in range(2):
            _lib.check(L.sscvae_test_gemm_splitk(_lib.ptr(A),K,_lib.ptr(B),K,M,N,K,_lib.ptr(Cm),N,S,None,s))
            torch.cuda.synchronize(); print("--", flush=True)
