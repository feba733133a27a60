"""Time one persistent decode-stack launch (256 scores) at a short and a long cache, CUDA events, 20 launches each."""
import sys, torch
sys.path.insert(0, ".")
from tests import parity
from scoreperformer_b200 import kernels as K
from scoreperformer_b200.decode import _StackWeights

model = parity.build_model(dropout=False, device="cuda").eval()
tr = model.perf_decoder.model.transformer
sw = _StackWeights(tr)
B, cap, S = 256, 1024, 64
kv = [torch.randn(B, cap, 128, device="cuda").bfloat16() for _ in range(sw.depth)]
plan = K.DecodeStackPlan(sw.layers, sw.w_ada, sw.b_ada, kv, B, S)
x = torch.randn(B, 256, device="cuda")
style = torch.randn(B, S, device="cuda")
for pos in (4, 255, 1023):
    pos_t = torch.tensor([pos], device="cuda")
    for _ in range(3):
        plan.step(x, style, None, pos_t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        plan.step(x, style, None, pos_t)
    e1.record()
    torch.cuda.synchronize()
    print(f"pos {pos}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per note-step", flush=True)
