"""Launch the dominant kernels of the C2 training step in isolation (for `ncu --set full`): FFN1 forward GEMM, attention fwd."""
import sys
import torch
sys.path.insert(0, ".")
from scoreperformer_b200 import kernels as K

N, D = 32768, 256
x = torch.randn(N, D, device="cuda").bfloat16()
w1 = torch.randn(2048, D, device="cuda").bfloat16()
b1 = torch.randn(2048, device="cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for _ in range(3):
    flush.zero_()
    K.gemm(x, w1, bias=b1)                       # FFN1 forward: [32768,256] x [2048,256]^T -> bf16 [32768,2048]
qkv = torch.randn(64 * 512, 384, device="cuda").bfloat16()
mask = torch.ones(64, 512, dtype=torch.bool, device="cuda")
ls = torch.log(torch.tensor([0.25, 0.0625, 0.015625, 0.0039], device="cuda"))
for impl in ("mma", "tcgen05"):
    for _ in range(2):
        flush.zero_()
        out, lse, aux = K.attention_fwd(qkv, mask, ls, 64, 512, 4, False, 0.1, 1, impl=impl)
        flush.zero_()
        K.attention_bwd(qkv, mask, ls, out, torch.randn_like(out), lse, torch.zeros(4, device="cuda"), 64, 512, 4, False, 0.1, 1, aux=aux)
torch.cuda.synchronize()
print("done")
