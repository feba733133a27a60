"""One forward + backward launch of the default attention at C2 (or `c4`) for an ncu capture:
   ncu --set full --clock-control none --import-source on -k regex:attn_ -o gpurun_out/X python tests/cuda/attn_one.py [c4] [causal] [drop]"""
import sys
import torch
sys.path.insert(0, ".")
from scoreperformer_b200 import kernels as K

args = sys.argv[1:]
B, T = (16, 2048) if "c4" in args else (64, 512)
causal, p, H = "causal" in args, 0.1 if "drop" in args else 0.0, 4
qkv = torch.randn(B * T, 384, device="cuda").bfloat16()
mask = torch.ones(B, T, dtype=torch.bool, device="cuda")
mask[:, T - T // 16:] = False
mask[0] = True
ls = torch.log(torch.tensor([0.25, 0.0625, 0.015625, 0.0039], device="cuda"))
for _ in range(2):
    out, lse, aux = K.attention_fwd(qkv, mask, ls, B, T, H, causal, p, 1)
    dout = torch.randn_like(out)
    delta = (dout.float().view(B, T, H, 64) * out.float().view(B, T, H, 64)).sum(-1).permute(0, 2, 1).contiguous()
    dls = torch.zeros(4, device="cuda")
    K.attention_bwd(qkv, mask, ls, out, dout, lse, dls, B, T, H, causal, p, 1, delta=delta, aux=aux)
torch.cuda.synchronize()
