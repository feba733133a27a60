"""CUPTI kernel durations (no launch latency) of the fused feed-forward kernels at C2 size, L2 flushed between launches."""
import sys
import torch
sys.path.insert(0, ".")
from scoreperformer_b200 import kernels as K
from torch.profiler import profile, ProfilerActivity

n, D, H = 32768, 256, 1024
xn = torch.randn(n, D, device="cuda").bfloat16()
w1 = (torch.randn(2 * H, D, device="cuda") / 16).bfloat16()
b1 = torch.randn(2 * H, device="cuda") * 0.1
w2 = (torch.randn(D, H, device="cuda") / 32).bfloat16()
w2t = K.transpose_bf16(w2)
resid = torch.randn(n, D, device="cuda")
dy = torch.randn(n, D, device="cuda").bfloat16()
db = torch.zeros(2 * H, device="cuda")
flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device="cuda")
cases = {"ffn_fwd save": lambda: K.ffn_fwd(xn, w1, b1, w2, resid, 0.1, 1, save=True),
         "ffn_fwd no-save": lambda: K.ffn_fwd(xn, w1, b1, w2, resid, 0.1, 1, save=False)}
out, u, h = K.ffn_fwd(xn, w1, b1, w2, resid, 0.1, 1, save=True)
cases["ffn_bwd"] = lambda: K.ffn_bwd(dy, w2t, w1, u, db, 0.1, 1, in_place=False)
for name, fn in cases.items():
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(10):
            flush.zero_()
            fn()
        torch.cuda.synchronize()
    d = [e.time_range.end - e.time_range.start for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "ffn_" in e.name]
    print(f"{name}: {sum(d) / len(d):.1f} us (min {min(d):.1f}, max {max(d):.1f})", flush=True)
