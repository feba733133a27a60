"""One GEMM shape of the C2 step in isolation for `ncu --set full`: python tests/cuda/profile_gemm.py {ffn1|ffn2|out|wgrad|dxn}"""
import sys
import torch
sys.path.insert(0, ".")
from scoreperformer_b200 import kernels as K

which = sys.argv[1] if len(sys.argv) > 1 else "ffn1"
M = 32768
bf = lambda *s: torch.randn(*s, device="cuda").bfloat16()
if which == "ffn1":
    a, w, kw = bf(M, 256), bf(2048, 256), dict(bias=torch.randn(2048, device="cuda"), out=torch.empty(M, 2048, dtype=torch.bfloat16, device="cuda"))
elif which == "ffn2":
    a, w, kw = bf(M, 1024), bf(256, 1024), dict(residual=torch.randn(M, 256, device="cuda"), out=torch.empty(M, 256, device="cuda"))
elif which == "out":
    a, w, kw = bf(M, 256), bf(256, 256), dict(residual=torch.randn(M, 256, device="cuda"), rowmask=torch.ones(M, dtype=torch.bool, device="cuda"),
                                              out=torch.empty(M, 256, device="cuda"))
elif which == "dxn":      # feed-forward backward: dxn[M,256] = du[M,2048] W1[2048,256]  (the largest GEMM share of the r02 step)
    a, w, kw = bf(M, 2048), bf(2048, 256), dict(trans_b=True, out=torch.empty(M, 256, dtype=torch.bfloat16, device="cuda"))
elif which == "wgrad":
    a, w, kw = bf(M, 2048), bf(M, 256), dict(trans_a=True, trans_b=True, split_k=0, out=torch.empty(2048, 256, device="cuda"))
for _ in range(3):
    K.gemm(a, w, **kw)
torch.cuda.synchronize()
print("done")
