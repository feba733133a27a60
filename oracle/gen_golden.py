"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the authoring container (where /root/reference exists):

    python oracle/gen_golden.py

It builds the reference ScorePerformer (default recipe) through oracle/ref_shim.py, overwrites the
weights with oracle/weights.fill_model_, switches every dropout to 0 (RNG parity is
not a goal), records the MMD prior samples the reference draws, and stores inputs-by-seed plus
outputs (losses, logits, hidden states, latents, selected gradients and all gradient norms).
The fixtures are small; the GPU box only ever reads the .npz files.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_shim  # noqa: E402
from weights import fill_model_  # noqa: E402
from scoreperformer_b200.synthetic import make_batch, SyntheticTokenizer  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

GRAD_KEYS = [  # full gradients stored for these (small or representative) tensors
    "perf_decoder.model.token_emb.embs.Velocity.index_weight",
    "perf_decoder.model.token_emb.embs.Velocity.value_layer.1.0.weight",
    "perf_decoder.model.token_emb.embs.Bar.value_layer.0.0.weight",
    "perf_decoder.model.token_emb.norm.weight",
    "perf_decoder.model.lm_head.norm.weight",
    "perf_decoder.model.emb_norm.bias",
    "perf_decoder.model.transformer.layers.0.1.to_k.weight",
    "perf_decoder.model.transformer.layers.0.1.rel_pos.learned_logslopes",
    "perf_decoder.model.transformer.layers.3.0.0.linear.bias",
    "perf_decoder.model.transformer.layers.7.1.ff.0.proj.bias",
    "perf_decoder.model.transformer.final_norm.linear.weight",
    "perf_encoder.transformer.layers.2.1.to_v.weight",
    "perf_encoder.transformer.layers.2.1.rel_pos.learned_logslopes",
    "perf_encoder.vae_head.mean.linear.weight",
    "perf_encoder.vae_head.bar_mean.linear.weight",
    "perf_encoder.vae_head.beat_mean.linear.weight",
    "perf_encoder.vae_head.onset_mean.linear.weight",
    "perf_encoder.vae_head.onset_mean.linear.bias",
    "score_encoder.transformer.layers.1.0.0.weight",
    "score_encoder.transformer.final_norm.bias",
    "score_encoder.token_emb.project_emb.bias",
    "classifiers.heads.dynamic/absolute.layers.0.weight",
    "classifiers.heads.articulation/tenuto.layers.0.bias",
]


def zero_dropouts(cfg):
    for stack in ("score_encoder", "perf_encoder", "perf_decoder"):
        cfg[stack]["transformer"]["attention"]["dropout"] = 0.0
        cfg[stack]["transformer"]["feed_forward"]["dropout"] = 0.0
    cfg["perf_encoder"]["latent_dropout"] = [0.0] * len(cfg["perf_encoder"]["latent_dim"])
    cfg["classifiers"]["classifier"]["dropout"] = 0.0
    return cfg


def build(seed_weights=0):
    ref_shim.install_stubs()
    from scoreperformer.models import ScorePerformer
    cfg = zero_dropouts(ref_shim.default_model_config())
    model = ScorePerformer.init(ref_shim._wrap(cfg))
    fill_model_(model, seed_weights)
    return model, cfg


class RecordRandn:
    """Capture the N(0,I) prior draws of MMDLoss.forward (mmd_transformer.py:519)."""

    def __init__(self):
        self.samples = []
        self._orig = torch.randn

    def __enter__(self):
        def randn(*a, **k):
            out = self._orig(*a, **k)
            self.samples.append(out.detach().clone())
            return out
        torch.randn = randn
        return self

    def __exit__(self, *exc):
        torch.randn = self._orig


def gen_train(name, B, T, seed):
    model, cfg = build()
    model.train()
    batch = make_batch(B, T, seed=seed)
    torch.manual_seed(99)
    with RecordRandn() as rec:
        out = model(**batch)
    out.loss.backward()
    g = {}
    sd = model.state_dict(keep_vars=True)
    norms = {}
    seen = set()
    for k, p in sd.items():
        if isinstance(p, torch.nn.Parameter) and p.grad is not None and p.data_ptr() not in seen:
            seen.add(p.data_ptr())
            norms[k] = float(p.grad.norm())
    arrays = {
        "B": B, "T": T, "seed": seed,
        "loss": out.loss.detach().numpy(),
        "loss_keys": np.array(list(out.losses.keys())),
        "loss_vals": np.array([float(v) for v in out.losses.values()], dtype=np.float64),
        "score_hidden": out.score_encoder.hidden_state.detach().numpy(),
        "perf_hidden": out.perf_encoder.hidden_state.detach().numpy(),
        "embeddings": out.perf_encoder.embeddings.detach().numpy(),
        "dec_hidden": out.perf_decoder.hidden_state.detach().numpy(),
        "grad_norm_keys": np.array(list(norms.keys())),
        "grad_norm_vals": np.array(list(norms.values()), dtype=np.float64),
    }
    for i, z in enumerate(rec.samples):
        arrays[f"z{i}"] = z.numpy()
    for i, lat in enumerate(out.perf_encoder.latents):
        arrays[f"latents{i}"] = lat.detach().numpy()
    for key in ("Velocity", "Tempo", "RelOnsetDev", "RelPerfDuration", "Bar", "NotesInOnset"):
        arrays[f"logits/{key}"] = out.perf_decoder.logits[key].detach().numpy().astype(np.float32)
    for key, lg in out.classifiers.logits.items():
        arrays[f"clf_logits/{key}"] = lg.detach().numpy()
    for k in GRAD_KEYS:
        arrays[f"grad/{k}"] = sd[k].grad.detach().numpy()
    # the reference evaluator on these outputs (evaluator.py:48-106), recipe settings (base.yaml:195-198) and their complement
    from scoreperformer.models.scoreperformer.evaluator import ScorePerformerEvaluator
    recipe_ignore = ["Bar", "Position", "Pitch", "Duration", "TimeSig", "PositionShift", "NotesInOnset", "PositionInOnset"]
    for tag, kw in (("recipe", dict(weighted_distance=True, ignore_keys=recipe_ignore)), ("plain", dict(weighted_distance=False))):
        ev = ScorePerformerEvaluator(model, tokenizer=SyntheticTokenizer(), **kw)
        metrics = ev({"labels": batch["labels"]}, out)
        arrays[f"eval/{tag}/keys"] = np.array(list(metrics.keys()))
        arrays[f"eval/{tag}/vals"] = np.array([float(v) for v in metrics.values()], dtype=np.float64)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name), **arrays)
    print(name, "loss", float(out.loss), "n_z", len(rec.samples),
          {k: round(float(v), 5) for k, v in out.losses.items()})


def gen_init(name, seed=23):
    """Constructor parity: key order, shapes and a checksum of every tensor the UNMODIFIED reference constructor produces under
    torch.manual_seed(seed) (same registration order => bit-identical initialisation)."""
    ref_shim.install_stubs()
    from scoreperformer.models import ScorePerformer
    torch.manual_seed(seed)
    model = ScorePerformer.init(ref_shim._wrap(ref_shim.default_model_config()))
    sd = model.state_dict()
    keys = list(sd.keys())
    np.savez_compressed(
        os.path.join(GOLDEN_DIR, name), seed=seed, keys=np.array(keys),
        shapes=np.array(["x".join(map(str, sd[k].shape)) for k in keys]),
        sums=np.array([float(sd[k].double().sum()) for k in keys], dtype=np.float64),
        abs_sums=np.array([float(sd[k].double().abs().sum()) for k in keys], dtype=np.float64),
        first=np.array([float(sd[k].reshape(-1)[0]) if sd[k].numel() else 0.0 for k in keys], dtype=np.float64))
    print(name, len(keys), "tensors")


def gen_collator(name, seed=11):
    """MixedLM masking of the UNMODIFIED collator (data/collators/performance.py:239-255) with the recipe's settings
    (base.yaml:62-65) on a seeded token tensor that contains every special token and padded tails."""
    ref_shim.install_stubs()
    from scoreperformer.data.collators.performance import MixedLMPerformanceCollator
    g = torch.Generator().manual_seed(seed)
    B, T, F = 5, 37, 12
    seq = torch.stack([torch.randint(0, v, (B, T), generator=g) for v in (260, 132, 92, 132, 133, 125, 26, 69, 16, 16, 165, 85)], dim=-1)
    seq[torch.rand(B, T, F, generator=g) < 0.15] = 2          # sprinkle SOS / EOS / MASK / PAD
    seq[torch.rand(B, T, F, generator=g) < 0.05] = 3
    seq[torch.rand(B, T, F, generator=g) < 0.05] = 1
    lengths = torch.tensor([37, 30, 1, 36, 18])
    seq = seq * (torch.arange(T)[None] < lengths[:, None])[..., None]
    out = {"seq": seq.numpy(), "lengths": lengths.numpy()}
    for tag, kw in (("recipe", dict(mask_ignore_token_ids=[0, 1, 2, 3], mask_ignore_token_dims=[0, 1, 2, 4, 6, 7, 8, 9])),
                    ("all_dims", dict(mask_ignore_token_ids=[0, 3], mask_ignore_token_dims=[], label_pad_ignored_dims=False)),
                    ("keep_labels", dict(mask_ignore_token_ids=[0, 1, 2, 3], mask_ignore_token_dims=[0, 5], label_pad_ignored_dims=False))):
        col = MixedLMPerformanceCollator(**kw)
        masked, labels = col.mask_sequence(seq)
        out[f"{tag}/masked"], out[f"{tag}/labels"] = masked.numpy(), labels.numpy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, name), **out)
    print(name, "ok")


def gen_render(name, T, seed):
    """Eval-mode encoders + cached greedy unmask_tokens for one score (generators.py:230-240).  Every greedy decision is recorded
    together with the gap between its two largest logits, so a test can demand exact tokens wherever the decision is not a
    numerical coin toss."""
    from scoreperformer.modules.sampling import top_k
    gaps, scales = [], []

    def recording_top_k(logits, **kw):
        top2 = logits.float().topk(2, dim=-1).values
        gaps.append(float((top2[..., 0] - top2[..., 1]).min()))
        finite = logits.float()[torch.isfinite(logits)]
        scales.append(float(finite.abs().max()))           # banned tokens are -inf: the scale of the real logits
        return top_k(logits, **kw)

    model, cfg = build()
    model.eval()
    batch = make_batch(1, T, seed=seed, full_length=True, deadpan_last=False)
    with torch.inference_mode():
        enc = model.forward_encoders(perf=batch["perf"], perf_mask=batch["perf_mask"], score=batch["score"],
                                     score_mask=batch["score_mask"], bars=batch["bars"], beats=batch["beats"],
                                     onsets=batch["onsets"], deadpan_mask=batch["deadpan_mask"], compute_loss=False)
        tokens = batch["masked_perf"].clone()       # fields {3,5,10,11} are MASK from note 0 on
        tokens[:, 0] = batch["perf"][:, 0]          # first note is given (idx-1 must exist)
        outs = {}
        for cached in (True, False):
            res = model.perf_decoder.unmask_tokens(
                tokens, batch["masked_perf"], filter_logits_fn=recording_top_k if cached else top_k, filter_kwargs={"k": 1},
                caches=None if cached else None, return_caches=False, disable_tqdm=True,
                context=enc.score_embeddings, style_embeddings=enc.perf_embeddings)
            outs[cached] = res
    np.savez_compressed(
        os.path.join(GOLDEN_DIR, name), T=T, seed=seed,
        score_embeddings=enc.score_embeddings.numpy(), perf_embeddings=enc.perf_embeddings.numpy(),
        tokens_in=tokens.numpy(), tokens_out=outs[True].numpy(),
        top2_gaps=np.array(gaps, dtype=np.float32),        # one per sampled field, in decoding order (note-major, then field)
        logit_scales=np.array(scales, dtype=np.float32))   # max |logit| of the same decisions
    print(name, "rendered", T, "notes; changed fields:", int((outs[True] != tokens).sum()), "decisions:", len(gaps),
          "min top-2 gap:", min(gaps) if gaps else None)


def gen_known_answers(name):
    """Appendix D vectors, recomputed from the reference's pure functions."""
    ref_shim.install_stubs()
    from scoreperformer.models.scoreperformer.mmd_transformer import MMDTupleTransformer, MMDLoss
    from scoreperformer.modules.transformer.embeddings import ALiBiPositionalBias
    from scoreperformer.modules.sampling import top_k
    from scoreperformer.models.classifiers.model import MultiHeadEmbeddingClassifier
    emb = torch.tensor([[[1., 2], [3, 4], [5, 6], [7, 8], [0, 0], [0, 0]]])
    seg = torch.tensor([[4, 4, 5, 6, 0, 0]])
    lat = MMDTupleTransformer._embeddings_to_latents(emb, "bar_mean", segments=seg)
    back = MMDTupleTransformer._latents_to_embeddings(lat, 6, "bar_mean", segments=seg)
    x = torch.tensor([[0., 0], [1, 0]])
    y = torch.tensor([[1., 1], [0, 2], [2, 2]])
    np.savez_compressed(
        os.path.join(GOLDEN_DIR, name),
        seg_emb=emb.numpy(), seg_ids=seg.numpy(), seg_latents=lat.numpy(), seg_back=back.numpy(),
        mmd_x=x.numpy(), mmd_y=y.numpy(), mmd_kernel=MMDLoss.gaussian_kernel(x, y).numpy(),
        mmd_value=MMDLoss.compute_mmd(x, y).numpy(),
        alibi=ALiBiPositionalBias(heads=4, total_heads=4, symmetric=True)(2, 4, 2).numpy(),
        topk=top_k(torch.arange(10.)[None]).numpy(),
        class_weights=np.array(MultiHeadEmbeddingClassifier._class_weights([0.9, 0.09, 0.01])))
    print(name, "ok")


if __name__ == "__main__":
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    gen_known_answers("known_answers.npz")
    gen_train("train_b2_t48.npz", 2, 48, seed=1234)
    gen_train("train_b3_t33.npz", 3, 33, seed=77)     # ragged: T-1 = 32, odd sizes
    gen_render("render_t24.npz", 24, seed=5)
    gen_render("render_t256.npz", 256, seed=9)
    gen_init("init_seed23.npz", seed=23)
    gen_collator("collator_mixlm.npz")
