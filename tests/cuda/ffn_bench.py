"""Fused feed-forward forward vs the unfused chain at C2 size (32768 rows), CUDA-event timed with an L2 flush between launches."""
import sys
import torch
sys.path.insert(0, ".")
from scoreperformer_b200 import kernels as K

n, D, H = 32768, 256, 1024
flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device="cuda")
xn = torch.randn(n, D, device="cuda").bfloat16()
w1 = (torch.randn(2 * H, D, device="cuda") / 16).bfloat16()
b1 = torch.randn(2 * H, device="cuda") * 0.1
w2 = (torch.randn(D, H, device="cuda") / 32).bfloat16()
resid = torch.randn(n, D, device="cuda")


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e3


def unfused(p):
    u = K.gemm(xn, w1, bias=b1, out_dtype=torch.bfloat16)
    h = K.glu_fwd(u, p, 1)
    return K.gemm(h, w2, residual=resid, out_dtype=torch.float32)


flops = 2.0 * n * D * (2 * H + H)
for p in (0.0, 0.1):
    for save in (True, False):
        us = timed(lambda: K.ffn_fwd(xn, w1, b1, w2, resid, p, 1, save=save))
        print(f"fused   drop={p} save_u_h={save}: {us:7.1f} us  {flops / us / 1e6:6.0f} TF/s", flush=True)
    us = timed(lambda: unfused(p))
    print(f"unfused drop={p}: {us:7.1f} us  {flops / us / 1e6:6.0f} TF/s", flush=True)
