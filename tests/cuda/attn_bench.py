"""Attention forward / backward in isolation at C2 (64 x 512) and C4 (16 x 2048), both implementations, CUDA-event timed with an L2
flush between launches.  Prints one line per case: microseconds and algorithmic TFLOP/s (4*B*H*T*T*64 fwd, x2.5 bwd, causal 1/2)."""
import sys
import torch
sys.path.insert(0, ".")
from scoreperformer_b200 import kernels as K

H = 4
flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1e3


for B, T in ((64, 512), (16, 2048)):
    qkv = torch.randn(B * T, 384, device="cuda").bfloat16()
    mask = torch.ones(B, T, dtype=torch.bool, device="cuda")
    mask[:, T - T // 16:] = False
    mask[0] = True
    ls = torch.log(torch.tensor([0.25, 0.0625, 0.015625, 0.0039], device="cuda"))
    for impl in sys.argv[1:] or ("tcgen05", "mma"):
        for causal in (False, True):
            for p in (0.0, 0.1):
                flops = 4.0 * B * H * T * T * 64 * (0.5 if causal else 1.0)
                us = timed(lambda: K.attention_fwd(qkv, mask, ls, B, T, H, causal, p, 1, impl=impl))
                out, lse, aux = K.attention_fwd(qkv, mask, ls, B, T, H, causal, p, 1, impl=impl)
                dout = torch.randn_like(out)
                delta = (dout.float().view(B, T, H, 64) * out.float().view(B, T, H, 64)).sum(-1).permute(0, 2, 1).contiguous()
                dls = torch.zeros(4, device="cuda")
                usb = timed(lambda: K.attention_bwd(qkv, mask, ls, out, dout, lse, dls, B, T, H, causal, p, 1, delta=delta, aux=aux))
                print(f"{impl:8s} B={B} T={T} {'causal' if causal else 'full  '} drop={p}: fwd {us:7.1f} us {flops / us / 1e6:6.0f} TF/s | "
                      f"bwd {usb:7.1f} us {2.5 * flops / usb / 1e6:6.0f} TF/s", flush=True)
