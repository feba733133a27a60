"""Tuple-embedding backward (stats + table scatter) at the C2 size in isolation, for ncu."""
import sys
import torch
sys.path.insert(0, ".")
from scoreperformer_b200 import kernels as K
from scoreperformer_b200.synthetic import make_batch

sizes = [260, 132, 92, 132, 133, 125, 26, 69, 16, 16, 165, 85]
b = make_batch(64, 512, seed=0)
tokens = b["perf"].cuda().view(-1, 12).contiguous()
n = tokens.shape[0]
table = torch.randn(sum(sizes), 128, device="cuda")
w, bias = torch.ones(1536, device="cuda"), torch.zeros(1536, device="cuda")
out, mean, rstd = K.embed_ln_fwd(tokens, table, sizes, w, bias)
dy = torch.randn(n, 1536, device="cuda").bfloat16()
for _ in range(3):
    dtable, dw, db = torch.zeros_like(table), torch.zeros_like(w), torch.zeros_like(bias)
    K.embed_ln_bwd(dy, tokens, table, sizes, w, mean, rstd, dtable, dw, db)
torch.cuda.synchronize()
print("done")
