"""One launch of the fused feed-forward backward at C2 size for an ncu capture: python tests/cuda/ffn_bwd_one.py [drop]"""
import sys
import torch
sys.path.insert(0, ".")
from scoreperformer_b200 import kernels as K

n, D, H = 32768, 256, 1024
dy = torch.randn(n, D, device="cuda").bfloat16()
u = torch.randn(n, 2 * H, device="cuda").bfloat16()
w1 = (torch.randn(2 * H, D, device="cuda") / 16).bfloat16()
w2t = K.transpose_bf16((torch.randn(D, H, device="cuda") / 32).bfloat16())
db = torch.zeros(2 * H, device="cuda")
for _ in range(2):
    K.ffn_bwd(dy, w2t, w1, u, db, 0.1 if "drop" in sys.argv else 0.0, 1, in_place=True)
torch.cuda.synchronize()
