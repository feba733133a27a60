"""One launch of the fused feed-forward forward at C2 size for an ncu capture: python tests/cuda/ffn_one.py [save] [drop]"""
import sys
import torch
sys.path.insert(0, ".")
from scoreperformer_b200 import kernels as K

n, D, H = 32768, 256, 1024
xn = torch.randn(n, D, device="cuda").bfloat16()
w1 = (torch.randn(2 * H, D, device="cuda") / 16).bfloat16()
b1 = torch.randn(2 * H, device="cuda") * 0.1
w2 = (torch.randn(D, H, device="cuda") / 32).bfloat16()
resid = torch.randn(n, D, device="cuda")
for _ in range(2):
    K.ffn_fwd(xn, w1, b1, w2, resid, 0.1 if "drop" in sys.argv else 0.0, 1, save="save" in sys.argv)
torch.cuda.synchronize()
