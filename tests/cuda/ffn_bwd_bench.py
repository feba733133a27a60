"""Fused feed-forward backward (data path) vs the unfused chain: correctness at a few sizes, then CUDA-event timing at C2 size
(32768 rows) with an L2 flush between launches."""
import sys
import torch
sys.path.insert(0, ".")
from scoreperformer_b200 import kernels as K

D, H = 256, 1024
BF16 = torch.bfloat16


def rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp(min=1e-6))


def unfused(dy, w2, w1, u, db, p, seed):
    dh = K.gemm(dy, w2, trans_b=True, out_dtype=BF16)
    du = K.glu_bwd(dh, u, db, p, seed)
    return K.gemm(du, w1, trans_b=True, out_dtype=BF16), du


torch.manual_seed(3)
w1 = (torch.randn(2 * H, D, device="cuda") / 16).bfloat16()
w2 = (torch.randn(D, H, device="cuda") / 32).bfloat16()
w2t = K.transpose_bf16(w2)
assert torch.equal(w2t, w2.t().contiguous())
for n in (256, 333, 4099, 32768):
    for p in (0.0, 0.25):
        dy = torch.randn(n, D, device="cuda").bfloat16()
        u = torch.randn(n, 2 * H, device="cuda").bfloat16()
        db_a = torch.zeros(2 * H, device="cuda")
        db_b = torch.zeros(2 * H, device="cuda")
        dxn_r, du_r = unfused(dy, w2, w1, u, db_a, p, 9)
        dxn, du = K.ffn_bwd(dy, w2t, w1, u, db_b, p, 9, in_place=False)
        torch.cuda.synchronize()
        print(f"n={n} p={p}: du {rel(du, du_r):.2e}  dxn {rel(dxn, dxn_r):.2e}  db1 {rel(db_b, db_a):.2e}  "
              f"mask equal {bool(torch.equal(du == 0, du_r == 0))}", flush=True)
        u2 = u.clone()
        dxn2, du2 = K.ffn_bwd(dy, w2t, w1, u2, None, p, 9, in_place=True)
        assert du2.data_ptr() == u2.data_ptr() and torch.equal(du2, du) and torch.equal(dxn2, dxn)

n = 32768
flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device="cuda")
dy = torch.randn(n, D, device="cuda").bfloat16()
u = torch.randn(n, 2 * H, device="cuda").bfloat16()
db = torch.zeros(2 * H, device="cuda")


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


flops = 2.0 * n * D * (H + 2 * H)
for p in (0.0, 0.1):
    for inp in (False, True):
        us = timed(lambda: K.ffn_bwd(dy, w2t, w1, u, db, p, 1, in_place=inp))
        print(f"fused   drop={p} in_place={inp}: {us:7.1f} us  {flops / us / 1e6:6.0f} TF/s", flush=True)
    us = timed(lambda: unfused(dy, w2, w1, u, db, p, 1))
    print(f"unfused drop={p}: {us:7.1f} us  {flops / us / 1e6:6.0f} TF/s", flush=True)
