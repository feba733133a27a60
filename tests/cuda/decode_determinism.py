import sys, torch
sys.path.insert(0, ".")
from tests import parity
from scoreperformer_b200 import kernels as K
from scoreperformer_b200.decode import _StackWeights
model = parity.build_model(dropout=False, device="cuda").eval()
tr = model.perf_decoder.model.transformer
sw = _StackWeights(tr)
B, cap, S = 3, 64, 64
torch.manual_seed(0)
kv0 = [torch.randn(B, cap, 128, device="cuda").bfloat16() for _ in range(sw.depth)]
x = torch.randn(B, 256, device="cuda")
style = torch.randn(B, S, device="cuda")
for pos in (0, 3, 23, 40):
    outs = []
    for rep in range(6):
        kv = [t.clone() for t in kv0]
        plan = K.DecodeStackPlan(sw.layers, sw.w_ada, sw.b_ada, kv, B, S)
        pos_t = torch.tensor([pos], device="cuda")
        out = plan.step(x, style, None, pos_t)
        torch.cuda.synchronize()
        outs.append(out.clone() if isinstance(out, torch.Tensor) else out[0].clone())
    same = [bool(torch.equal(outs[0], o)) for o in outs[1:]]
    print(pos, same, float((outs[0] - outs[-1]).abs().max()))
