"""Time every GEMM shape of the C2 training step (M = 64 x 512 note-tuples) in isolation: us, TFLOP/s, algorithmic GB/s.
Usage: python tests/cuda/gemm_shapes.py [BN override via SPB_GEMM_BN]"""
import sys
import torch
sys.path.insert(0, ".")
from scoreperformer_b200 import kernels as K

M = 32768
BF, F32 = torch.bfloat16, torch.float32
dev = "cuda"


def t(fn, iters=20):
    """GPU time per launch: `iters` launches captured in one CUDA graph (no CPU launch floor), replayed 3 times."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * iters) * 1e3


def case(name, m, n, k, ta=False, tb=False, out=BF, bias=False, residual=False, rowmask=False, split_k=1):
    a = torch.randn((k, m) if ta else (m, k), device=dev).to(BF)
    b = torch.randn((k, n) if tb else (n, k), device=dev).to(BF)
    bi = torch.randn(n, device=dev) if bias else None
    res = torch.randn(m, n, device=dev) if residual else None
    rm = torch.ones(m, dtype=torch.bool, device=dev) if rowmask else None
    c = torch.empty(m, n, dtype=out, device=dev)
    us = t(lambda: K.gemm(a, b, trans_a=ta, trans_b=tb, bias=bi, residual=res, rowmask=rm, out=c, split_k=split_k))
    flops = 2.0 * m * n * k
    byts = (m * k + n * k) * 2 + m * n * (4 if out == F32 else 2) + (m * n * 4 if residual else 0)
    print(f"{name:34s} M={m:6d} N={n:5d} K={k:6d}  {us:8.1f} us  {flops / us / 1e6:7.1f} TF/s  {byts / us / 1e3:7.0f} GB/s")


case("qkv fwd", M, 384, 256)
case("attn out fwd (+res,mask,f32)", M, 256, 256, out=F32, residual=True, rowmask=True)
case("ffn1 fwd (+bias)", M, 2048, 256, bias=True)
case("ffn2 fwd (+res,f32)", M, 256, 1024, out=F32, residual=True)
case("embed proj fwd K=1536", M, 256, 1536, bias=True, out=F32)
case("head proj fwd N=1536", M, 1536, 256, tb=True)
case("dgrad ffn2  (dh)", M, 1024, 256, tb=True)
case("dgrad ffn1  (dxn)", M, 256, 2048, tb=True)
case("dgrad qkv", M, 256, 384, tb=True)
case("dgrad out", M, 256, 256, tb=True)
case("wgrad ffn1 [2048x256] K=M", 2048, 256, M, ta=True, tb=True, out=F32, split_k=0)
case("wgrad ffn2 [256x1024] K=M", 256, 1024, M, ta=True, tb=True, out=F32, split_k=0)
case("wgrad qkv [384x256] K=M", 384, 256, M, ta=True, tb=True, out=F32, split_k=0)
case("wgrad out [256x256] K=M", 256, 256, M, ta=True, tb=True, out=F32, split_k=0)
