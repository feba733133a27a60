import sys, torch, os
sys.path.insert(0, '.')
from scoreperformer_b200 import kernels as K
B,T,H=64,512,4
qkv = torch.randn(B*T, 384, device='cuda').bfloat16()
mask = torch.ones(B,T,dtype=torch.bool,device='cuda'); mask[:, 480:] = False; mask[0]=True
ls = torch.log(torch.tensor([0.25,0.0625,0.015625,0.0039],device='cuda'))
for impl in ('mma','tcgen05'):
    K.ATTENTION_FWD_IMPL = impl
    for causal in (False, True):
        for p in (0.0, 0.1):
            for _ in range(3): K.attention_fwd(qkv, mask, ls, B, T, H, causal, p, 1)
            torch.cuda.synchronize()
            e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): K.attention_fwd(qkv, mask, ls, B, T, H, causal, p, 1)
            e1.record(); torch.cuda.synchronize()
            print(impl, 'causal' if causal else 'full', 'drop', p, '%.1f us' % (e0.elapsed_time(e1)/20*1e3))
K.ATTENTION_FWD_IMPL = 'mma'
for causal in (False, True):
    for p in (0.0, 0.1):
        out, lse = K.attention_fwd(qkv, mask, ls, B, T, H, causal, p, 1)
        dout = torch.randn_like(out)
        dls = torch.zeros(4, device='cuda')
        for _ in range(3): K.attention_bwd(qkv, mask, ls, out, dout, lse, dls, B, T, H, causal, p, 1)
        torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): K.attention_bwd(qkv, mask, ls, out, dout, lse, dls, B, T, H, causal, p, 1)
        e1.record(); torch.cuda.synchronize()
        print('bwd', 'causal' if causal else 'full', 'drop', p, '%.1f us' % (e0.elapsed_time(e1)/20*1e3))
