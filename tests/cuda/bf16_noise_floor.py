"""How far does the UNMODIFIED reference drift from itself when it computes in bf16?  Same weights (oracle/weights.fill_model_),
same batch, dropouts off: fp32 eager vs `torch.autocast(bfloat16)` eager on this GPU, with the deviation measures of
tests/parity.py (max abs error / max abs reference, and the rms ratio).  This is the noise floor any bf16 implementation of the
model lives on; profiles/r02_bf16_noise_floor.txt keeps the output.   python tests/cuda/bf16_noise_floor.py [B T]"""
import os
import sys
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
warnings.filterwarnings("ignore")
for cand in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
    if os.path.isdir(os.path.join(cand, "scoreperformer")):
        os.environ["SPB200_REFERENCE_ROOT"] = cand
        break
import ref_shim  # noqa: E402
from gen_golden import zero_dropouts  # noqa: E402
from weights import fill_model_  # noqa: E402
from scoreperformer_b200.synthetic import make_batch  # noqa: E402

B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4, 256)
ref_shim.install_stubs()
from scoreperformer.models import ScorePerformer  # noqa: E402

cfg = zero_dropouts(ref_shim.default_model_config())
model = ScorePerformer.init(ref_shim._wrap(cfg))
fill_model_(model, 0)
model = model.cuda().train()
batch = {k: v.cuda() for k, v in make_batch(B, T, seed=1234).items()}
torch.backends.cuda.matmul.allow_tf32 = False


def run(autocast: bool):
    torch.manual_seed(99)           # same MMD prior samples
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        return model(**batch)


a, b = run(False), run(True)


def dev(x, y):
    x, y = x.detach().float(), y.detach().float()
    return float((y - x).abs().max() / x.abs().max()), float((y - x).pow(2).mean().sqrt() / x.pow(2).mean().sqrt())


rows = {n: dev(getattr(a, m).hidden_state, getattr(b, m).hidden_state) for n, m in (("score_hidden", "score_encoder"), ("perf_hidden", "perf_encoder"), ("dec_hidden", "perf_decoder"))}
rows["embeddings"] = dev(a.perf_encoder.embeddings, b.perf_encoder.embeddings)
lg = {k: dev(v, b.perf_decoder.logits[k]) for k, v in a.perf_decoder.logits.items()}
worst = max(lg, key=lambda k: lg[k][0])
print(f"reference bf16-autocast vs reference fp32, B={B} T={T}  (max-rel / rms-rel)")
print(" ".join(f"{k}={v[0]:.4f}/{v[1]:.4f}" for k, v in rows.items()))
print(f"logits worst {worst}={lg[worst][0]:.4f}/{lg[worst][1]:.4f}  mean max-rel {sum(v[0] for v in lg.values()) / len(lg):.4f} "
      f"mean rms-rel {sum(v[1] for v in lg.values()) / len(lg):.4f}")
print("per field max-rel:", " ".join(f"{k}={v[0]:.4f}" for k, v in lg.items()))
print("losses fp32 / bf16:", " ".join(f"{k}={float(a.losses[k]):.4f}/{float(b.losses[k]):.4f}" for k in list(a.losses)[:4]))
