import sys, torch
sys.path.insert(0, ".")
from scoreperformer_b200 import kernels as K
from torch.profiler import profile, ProfilerActivity
torch.manual_seed(0)
sizes = [260, 132, 92, 132, 133, 125, 26, 69, 16, 16, 165, 85]
n = 32768
table = torch.randn(sum(sizes), 128, device="cuda")
tokens = torch.stack([torch.randint(0, v, (n,), device="cuda") for v in sizes], dim=-1)
w, b = torch.randn(1536, device="cuda") * 0.2 + 1, torch.randn(1536, device="cuda") * 0.1
x16, mean, rstd = K.embed_ln_fwd(tokens, table, sizes, w, b)
dy = torch.randn(n, 1536, device="cuda").bfloat16()
dtable = torch.zeros_like(table); dw = torch.zeros(1536, device="cuda"); db = torch.zeros(1536, device="cuda")
flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for _ in range(3): K.embed_ln_bwd(dy, tokens, table, sizes, w, mean, rstd, dtable, dw, db)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(8):
        flush.zero_(); K.embed_ln_bwd(dy, tokens, table, sizes, w, mean, rstd, dtable, dw, db)
    torch.cuda.synchronize()
import collections
agg = collections.defaultdict(list)
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA and "embed" in e.name: agg[e.name[:60]].append(e.time_range.end - e.time_range.start)
for k, v in agg.items(): print(k, f"{sum(v)/len(v):.1f} us")
