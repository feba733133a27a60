"""Print the forward deviations (max abs error / max abs reference) of one training step against the fp32 oracle, for the current
SPB_* environment switches.  Diagnostic for the bf16 parity budget: python tests/cuda/parity_probe.py [B T]"""
import sys
import torch
sys.path.insert(0, ".")
from tests import parity

B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4, 256)
torch.manual_seed(1234)
model = parity.build_model(dropout=False, device="cuda")
batch = parity.make_batch(B, T, seed=1234)
z = [torch.randn(256, d) for d in (32, 20, 8, 4)]
ref, sd = parity.run_oracle_step(model, batch, z, device="cuda")
out = parity.run_product_step(model, batch, z, mmd_rows=ref["mmd_rows"])


def relerr(a, b):
    b = b.detach()
    return float((a.detach().float().to(b.device) - b).abs().max() / b.abs().max())


def rms(a, b):
    b = b.detach()
    return float((a.detach().float().to(b.device) - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt())


row = {n: (relerr(t, ref[n]), rms(t, ref[n])) for n, t in (("score_hidden", out.score_encoder.hidden_state), ("perf_hidden", out.perf_encoder.hidden_state),
                                      ("embeddings", out.perf_encoder.embeddings), ("dec_hidden", out.perf_decoder.hidden_state))}
lg = {k: (relerr(out.perf_decoder.logits[k], v), rms(out.perf_decoder.logits[k], v)) for k, v in ref["logits"].items()}
worst = max(lg, key=lambda k: lg[k][0])
print(" ".join(f"{k}={v[0]:.4f}/{v[1]:.4f}" for k, v in row.items()), f"| logits worst {worst}={lg[worst][0]:.4f}/{lg[worst][1]:.4f}",
      f"mean max-rel {sum(v[0] for v in lg.values()) / len(lg):.4f} mean rms-rel {sum(v[1] for v in lg.values()) / len(lg):.4f}")
