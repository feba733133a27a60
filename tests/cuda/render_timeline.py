"""Kernel records (torch.profiler / CUPTI) of a short batched rendering: what a note-step consists of besides the persistent
decode-stack kernel.  python tests/cuda/render_timeline.py [notes]"""
import collections
import sys
import torch
sys.path.insert(0, ".")
from tests import parity
from scoreperformer_b200.decode import render_batch
from torch.profiler import profile, ProfilerActivity

T = int(sys.argv[1]) if len(sys.argv) > 1 else 200
B = 256
model = parity.build_model(dropout=False, device="cuda").eval()
batch = {k: v.cuda() for k, v in parity.make_batch(B, T, seed=3, full_length=True, deadpan_last=False).items()}
with torch.no_grad():
    enc = model.forward_encoders(perf=batch["perf"], perf_mask=batch["perf_mask"], score=batch["score"], score_mask=batch["score_mask"],
                                 bars=batch["bars"], beats=batch["beats"], onsets=batch["onsets"], deadpan_mask=batch["deadpan_mask"],
                                 compute_loss=False)
    tokens = batch["masked_perf"].clone()
    tokens[:, 0] = batch["perf"][:, 0]
    args = (model, tokens, batch["masked_perf"], enc.score_embeddings, enc.perf_embeddings)
    render_batch(*args)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        render_batch(*args)
        torch.cuda.synchronize()
ev = [(e.time_range.start, e.time_range.end - e.time_range.start, e.name) for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort()
span = ev[-1][0] + ev[-1][1] - ev[0][0]
agg = collections.defaultdict(lambda: [0, 0.0])
for s, d, n in ev:
    k = n.replace("(anonymous namespace)::", "").replace("void ", "").replace("at::native::", "")[:80]
    agg[k][0] += 1
    agg[k][1] += d
steps = T - 1
print(f"{B} scores x {T} notes: {span / 1e3:.1f} ms, {span / steps:.1f} us per note-step, {len(ev) / steps:.1f} kernels per step")
for k, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f"{d / steps:8.1f} us/step {c / steps:5.1f} x  {k}")
stack = [(s_, d) for s_, d, n in ev if "decode_stack_kernel" in n]
print(f"decode_stack launches: {len(stack)}; period between launches {(stack[-1][0] - stack[2][0]) / (len(stack) - 3):.1f} us "
      f"(steady state, graph replays); before the first one: {(stack[0][0] - ev[0][0]) / 1e3:.2f} ms")
busy = 0.0
cur_e = ev[0][0]
for s, d, n in ev:
    if s + d > cur_e:
        busy += s + d - max(s, cur_e)
        cur_e = s + d
print(f"busy {busy / steps:.1f} us per step, idle {(span - busy) / steps:.1f} us per step")
