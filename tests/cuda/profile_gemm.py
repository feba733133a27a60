"""FFN1-forward GEMM in isolation for `ncu --set full` (one launch per configuration after warm-up)."""
import os
import sys
import torch
sys.path.insert(0, ".")
from scoreperformer_b200 import kernels as K

N, D = 32768, 256
x = torch.randn(N, D, device="cuda").bfloat16()
w1 = torch.randn(2048, D, device="cuda").bfloat16()
b1 = torch.randn(2048, device="cuda")
out = torch.empty(N, 2048, dtype=torch.bfloat16, device="cuda")
for _ in range(3):
    K.gemm(x, w1, bias=b1, out=out)
torch.cuda.synchronize()
print("done")
