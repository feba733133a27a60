"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain PyTorch, fp32) of the reference hot path.

This file is the *oracle*: a functional re-statement of ScorePerformer's training / inference
forward for the default-recipe family, written against a `state_dict` with the reference's key
layout.  It exists so that the CUDA path can be checked on a box where
`/root/reference` is absent.  It is pinned to the real reference by `tests/golden/*.npz`, which
`oracle/gen_golden.py` produced by running the unmodified reference in the authoring container
(tests/test_oracle_golden.py re-checks that on every CPU run).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py --impl reference` / the `cpu_baseline`
leg may import this module.  The product package never does.

Each function cites the reference file:line (relative to /root/reference/scoreperformer) it follows.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F
from torch import Tensor

IGNORE_INDEX = -100


# ----------------------------------------------------------------------------- hyper-parameters
@dataclass
class OracleSpec:
    """Sizes of the default recipe (recipes/scoreperformer/base.yaml:68-192)."""
    num_tokens: Dict[str, int]
    num_score_tokens: Dict[str, int]
    num_classes: Dict[str, int]
    dim: int = 256
    emb_dim: int = 128
    heads: int = 4
    dim_head: int = 64
    depth_score: int = 2
    depth_perf: int = 4
    depth_dec: int = 4
    latent_dim: Sequence[int] = (32, 20, 8, 4)
    aggregate_mode: Sequence[str] = ("mean", "bar_mean", "beat_mean", "onset_mean")
    discrete_ids: Sequence[int] = (0, 1, 2, 3)
    mmd_loss_weight: float = 1.0
    clf_loss_weight: float = 1.0
    mmd_num_samples: int = 256
    mmd_max_latents: int = 4096
    deadpan_zero_latent: bool = True

    @property
    def style_dim(self) -> int:
        return int(sum(self.latent_dim))


# ----------------------------------------------------------------------------- a1: computed tables
def field_table(sd: Dict[str, Tensor], prefix: str, key: str, discrete_ids=(0, 1, 2, 3)) -> Tensor:
    """W_f = index rows {0,1,2,3} (others zero) + MLP(token_values) with those rows zeroed.

    modules/transformer/embeddings.py:124-143 (token_weight, value_weight), :199-211 (value MLP =
    Linear(1,E) -> Mish -> Linear(E,E)), :91-99 (weight = token_weight + value_weight).
    """
    p = f"{prefix}.embs.{key}."
    iw = sd[p + "index_weight"]
    ids = torch.as_tensor(list(discrete_ids), dtype=torch.long, device=iw.device)
    tok = torch.zeros_like(iw)
    tok[ids] = iw[ids]
    tv = sd[p + "token_values"].to(iw.dtype)
    h = F.mish(F.linear(tv, sd[p + "value_layer.0.0.weight"], sd[p + "value_layer.0.0.bias"]))
    val = F.linear(h, sd[p + "value_layer.1.0.weight"], sd[p + "value_layer.1.0.bias"])
    keep = torch.ones(val.shape[0], 1, dtype=val.dtype, device=val.device)
    keep[ids] = 0.0
    return tok + val * keep


def tuple_embed(sd, prefix: str, tokens: Tensor, keys: Sequence[str], spec: OracleSpec) -> Tensor:
    """gather F fields -> cat -> LayerNorm(E*F) -> Linear(E*F, dim).

    models/scoreperformer/embeddings.py:121-143.  F.embedding(padding_idx=0) only affects grads of
    row 0 (the row receives none); reproduced by detaching that row.
    """
    parts = []
    for i, key in enumerate(keys):
        w = field_table(sd, prefix, key, spec.discrete_ids)
        w = torch.cat([w[:1].detach(), w[1:]], dim=0)
        parts.append(w[tokens[..., i]])
    x = torch.cat(parts, dim=-1)
    x = F.layer_norm(x, (x.shape[-1],), sd[f"{prefix}.norm.weight"], sd[f"{prefix}.norm.bias"])
    return F.linear(x, sd[f"{prefix}.project_emb.weight"], sd[f"{prefix}.project_emb.bias"])


# ----------------------------------------------------------------------------- a5/a6: attention
def alibi_bias(i: int, j: int, k: int, slopes: Tensor) -> Tensor:
    """-slope_h * |col - (row + k)|; modules/transformer/embeddings.py:294-315 (symmetric)."""
    rows = torch.arange(k, i + k, device=slopes.device)
    cols = torch.arange(j, device=slopes.device)
    dist = -(cols[None, :] - rows[:, None]).abs().to(slopes.dtype)
    return slopes.view(-1, 1, 1) * dist[None]


def attention(sd, prefix: str, x: Tensor, mask: Optional[Tensor], causal: bool, spec: OracleSpec,
              cache_kv=None, attn_dropout_mask: Optional[Tensor] = None):
    """MQA attention with learned-slope ALiBi, key padding and causal masks.

    modules/transformer/attention.py:107-222 and attend.py:58-126.  Masked positions receive the
    additive fill -finfo.max//2 (attend.py:102-108); the output is zeroed at padded queries
    (attention.py:216-218).  Returns (out, k, v) where k, v include the cache (attention.py:155-156).
    """
    b, n, _ = x.shape
    h, dh = spec.heads, spec.dim_head
    q = F.linear(x, sd[f"{prefix}.to_q.weight"]).view(b, n, h, dh).transpose(1, 2)
    k = F.linear(x, sd[f"{prefix}.to_k.weight"])
    v = F.linear(x, sd[f"{prefix}.to_v.weight"])
    if cache_kv is not None:
        k = torch.cat([cache_kv[0], k], dim=1)
        v = torch.cat([cache_kv[1], v], dim=1)
    j = k.shape[1]
    slopes = sd[f"{prefix}.rel_pos.learned_logslopes"].exp().view(-1)
    bias = alibi_bias(n, j, j - n, slopes)[None].expand(b, -1, -1, -1)
    allowed = torch.ones(b, 1, n, j, dtype=torch.bool, device=x.device)
    if mask is not None:
        allowed = allowed & mask[:, None, None, :]
    if causal:
        allowed = allowed & ~torch.ones(n, j, dtype=torch.bool, device=x.device).triu(j - n + 1)
    fill = -torch.finfo(x.dtype).max // 2
    bias = bias.masked_fill(~allowed, fill)
    scores = torch.einsum("bhid,bjd->bhij", q, k) * (dh ** -0.5) + bias
    p = scores.softmax(dim=-1)
    if attn_dropout_mask is not None:
        p = p * attn_dropout_mask
    o = torch.einsum("bhij,bjd->bhid", p, v).transpose(1, 2).reshape(b, n, h * dh)
    o = F.linear(o, sd[f"{prefix}.to_out.weight"])
    if mask is not None:
        m = mask[:, -1:] if cache_kv is not None else mask
        o = o * m[..., None]
    return o, k, v


def feed_forward(sd, prefix: str, x: Tensor) -> Tensor:
    """GLU(SiLU) feed-forward; modules/transformer/feedforward.py:13-22,56-64."""
    u = F.linear(x, sd[f"{prefix}.ff.0.proj.weight"], sd[f"{prefix}.ff.0.proj.bias"])
    a, gate = u.chunk(2, dim=-1)
    return F.linear(a * F.silu(gate), sd[f"{prefix}.ff.3.weight"])


def _norm(sd, prefix: str, x: Tensor, style: Optional[Tensor]) -> Tensor:
    """nn.LayerNorm or AdaptiveLayerNorm (modules/layers.py:31-47)."""
    if style is None:
        return F.layer_norm(x, (x.shape[-1],), sd[f"{prefix}.weight"], sd[f"{prefix}.bias"])
    gb = F.linear(style, sd[f"{prefix}.linear.weight"], sd[f"{prefix}.linear.bias"])
    gamma, beta = gb.chunk(2, dim=-1)
    return gamma * F.layer_norm(x, (x.shape[-1],)) + beta


def transformer_stack(sd, prefix: str, x: Tensor, mask: Optional[Tensor], depth: int, causal: bool,
                      spec: OracleSpec, style: Optional[Tensor] = None, caches=None):
    """Pre-norm ('a','f') x depth stack + final norm; modules/transformer/transformer.py:139-232.

    Returns (out, hiddens, kvs): `hiddens` are the attention-layer inputs plus the final output
    (the KV/hidden cache layout of TransformerIntermediates, transformer.py:25-28).
    With `caches=(hiddens, kvs)` only the last position is computed (transformer.py:161-186).
    """
    has_cache = caches is not None
    if has_cache:
        c_h, c_kv = list(caches[0]), list(caches[1])
        x = x[:, -1:]
        style = style[:, -1:] if style is not None else None
    hiddens, kvs = [], []
    for layer in range(2 * depth):
        p = f"{prefix}.layers.{layer}"
        if layer % 2 == 0:
            if has_cache:
                x = torch.cat([c_h.pop(0), x], dim=1)
            hiddens.append(x)
            x = x[:, -1:] if has_cache else x
        res = x
        xn = _norm(sd, f"{p}.0.0", x, style)
        if layer % 2 == 0:
            o, k, v = attention(sd, f"{p}.1", xn, mask, causal, spec, cache_kv=c_kv.pop(0) if has_cache else None)
            kvs.append((k, v))
        else:
            o = feed_forward(sd, f"{p}.1", xn)
        x = o + res
    x = _norm(sd, f"{prefix}.final_norm", x, style)
    if has_cache:
        x = torch.cat([c_h.pop(0), x], dim=1)
    hiddens.append(x)
    return x, hiddens, kvs


# ----------------------------------------------------------------------------- a3: TupleTransformer
def encoder_forward(sd, name: str, tokens: Tensor, mask: Tensor, keys, depth: int, spec: OracleSpec) -> Tensor:
    """models/scoreperformer/transformer.py:146-222 for the two encoders (no context, no style)."""
    x = tuple_embed(sd, f"{name}.token_emb", tokens, keys, spec)
    x = F.layer_norm(x, (spec.dim,), sd[f"{name}.emb_norm.weight"], sd[f"{name}.emb_norm.bias"])
    out, _, _ = transformer_stack(sd, f"{name}.transformer", x, mask, depth, False, spec)
    return out


def decoder_embed(sd, seq: Tensor, seq_masked: Tensor, context: Tensor, spec: OracleSpec):
    """multi-seq 'post-cat' embedding + emb_norm + cat(context) + project_emb.

    models/scoreperformer/embeddings.py:245-255; transformer.py:171-185.  Returns (x, token_emb).
    """
    name = "perf_decoder.model"
    keys = list(spec.num_tokens)
    e1 = tuple_embed(sd, f"{name}.token_emb", seq, keys, spec)
    e2 = tuple_embed(sd, f"{name}.token_emb", seq_masked, keys, spec)
    te = F.linear(torch.cat([e1, e2], dim=-1), sd[f"{name}.token_emb.project_multiemb.weight"],
                  sd[f"{name}.token_emb.project_multiemb.bias"])
    x = F.layer_norm(te, (spec.dim,), sd[f"{name}.emb_norm.weight"], sd[f"{name}.emb_norm.bias"])
    x = torch.cat([x, context[:, :x.shape[1]]], dim=-1)
    x = F.linear(x, sd[f"{name}.project_emb.weight"], sd[f"{name}.project_emb.bias"])
    return x, te


def tied_head(sd, hidden: Tensor, spec: OracleSpec, keys: Optional[Sequence] = None) -> Dict[str, Tensor]:
    """LN_1536(h @ W_proj) split per field, times the computed table transposed.

    models/scoreperformer/embeddings.py:345-353 (TupleTokenTiedLMHead, reuse_projection).
    """
    name = "perf_decoder.model"
    e = hidden @ sd[f"{name}.token_emb.project_emb.weight"]
    e = F.layer_norm(e, (e.shape[-1],), sd[f"{name}.lm_head.norm.weight"], sd[f"{name}.lm_head.norm.bias"])
    out = {}
    for i, key in enumerate(spec.num_tokens):
        if keys is not None and i not in keys and key not in keys:
            continue
        w = field_table(sd, f"{name}.token_emb", key, spec.discrete_ids)
        out[key] = e[..., i * spec.emb_dim:(i + 1) * spec.emb_dim] @ w.t()
    return out


# ----------------------------------------------------------------------------- a8/a9: MMD-VAE levels
def segment_mean(x: Tensor, segments: Tensor):
    """Per-sample mean over runs of equal segment id -> [B, S, D], S = max id + 1.

    models/scoreperformer/mmd_transformer.py:330-340 (dense one-hot alignment + bmm, counts >= 1).
    Returned `counts` are the raw membership counts (bit-exact membership check).
    """
    b, t, d = x.shape
    s = int(segments.max()) + 1
    rows = torch.arange(b, device=x.device)[:, None].expand(b, t)
    sums = torch.zeros(b, s, d, dtype=x.dtype, device=x.device).index_put_((rows, segments), x, accumulate=True)
    counts = torch.zeros(b, s, dtype=torch.long, device=x.device).index_put_(
        (rows, segments), torch.ones(b, t, dtype=torch.long, device=x.device), accumulate=True)
    return sums / counts.clamp(min=1)[..., None].to(x.dtype), counts


def gaussian_kernel_mean(x: Tensor, y: Tensor) -> Tensor:
    """mean_ij exp(-||x_i - y_j||^2 / d^2); mmd_transformer.py:522-527 (note the double /d)."""
    d = x.shape[-1]
    d2 = (x[:, None, :] - y[None, :, :]).pow(2).mean(-1) / d
    return torch.exp(-d2).mean()


def mmd(z: Tensor, y: Tensor) -> Tensor:
    """mmd_transformer.py:529-534."""
    return gaussian_kernel_mean(z, z) + gaussian_kernel_mean(y, y) - 2 * gaussian_kernel_mean(z, y)


def perf_encoder_forward(sd, batch, spec: OracleSpec, z_prior: Optional[List[Tensor]], training: bool,
                         compute_loss: bool = True, mmd_perm: Optional[List[Optional[Tensor]]] = None):
    """MMDTupleTransformer.forward, hierarchical_with_context, dropouts off.

    models/scoreperformer/mmd_transformer.py:169-302 and _forward_latents :304-368.
    `z_prior[l]` is the injected N(0,I) sample [256, z_l] of MMDLoss.forward (:519).
    `mmd_perm[l]` injects the `randperm(n)[:max_num_latents]` draw of MMDLoss.forward (:515-517) for levels with more
    than `max_num_latents` valid latents; without it a fresh permutation is drawn.  The rows actually used are returned
    as `mmd_rows[l]` = (sample index, segment index) pairs, so a test can hand the SAME subsample to the CUDA path.
    """
    mask = batch["perf_mask"]
    hidden = encoder_forward(sd, "perf_encoder", batch["perf"], mask, list(spec.num_tokens), spec.depth_perf, spec)
    m3 = mask[..., None]
    out = hidden * m3
    b, t, _ = out.shape
    seg_of = {"bar_mean": batch.get("bars"), "beat_mean": batch.get("beats"), "onset_mean": batch.get("onsets")}
    latents, embs, losses, counts_all, mmd_rows = [], [], {}, [], []
    for lvl, (mode, zl) in enumerate(zip(spec.aggregate_mode, spec.latent_dim)):
        w = sd[f"perf_encoder.vae_head.{mode}.linear.weight"]
        bias = sd[f"perf_encoder.vae_head.{mode}.linear.bias"]
        if mode == "mean":
            pooled = (out.sum(dim=1) / m3.sum(dim=1))[:, None]                      # :325-327
            lmask = torch.ones(b, 1, dtype=torch.bool, device=out.device)
            counts_all.append(mask.sum(1, keepdim=True))
        else:
            pooled, counts = segment_mean(out, seg_of[mode])                         # :330-340
            lmask = torch.all(pooled != 0.0, dim=-1)                                 # :342
            counts_all.append(counts)
        lat = F.linear(pooled, w, bias) * lmask[..., None]                           # :346-347
        if mode == "mean":
            emb = lat.expand(-1, t, -1)                                              # :356-359
        else:
            emb = lat[torch.arange(b, device=lat.device)[:, None].expand(b, t), seg_of[mode]]           # :362-364
        emb = emb * m3                                                               # :366
        latents.append(lat)
        embs.append(emb)
        out = torch.cat([out, emb], dim=-1)                                          # :259-261
        if compute_loss:
            y = lat[lmask]
            rows = None
            if y.shape[0] > spec.mmd_max_latents:                                    # :515-517
                perm = mmd_perm[lvl] if mmd_perm is not None and mmd_perm[lvl] is not None else \
                    torch.randperm(y.shape[0], device=y.device)[:spec.mmd_max_latents]
                y = y[perm]
                rows = torch.nonzero(lmask)[perm]                                    # [max, 2] = (sample, segment)
            mmd_rows.append(rows)
            losses[f"MMD/{mode}"] = spec.mmd_loss_weight * mmd(z_prior[lvl].to(y.dtype), y)   # :266, :519-520
            if spec.deadpan_zero_latent:
                dp = lat[batch["deadpan_mask"][:, None] & lmask]                     # :268-273
                if bool(torch.any(dp != 0)):
                    losses[f"MMD/{mode}/deadpan"] = F.mse_loss(dp, torch.zeros_like(dp))
    embeddings = torch.cat(embs, dim=-1) * m3                                        # :275-278
    loss = None
    if compute_loss:
        loss = sum(losses.values())
        losses["MMD"] = loss
    return dict(hidden_state=hidden, latents=latents, embeddings=embeddings, full_embeddings=embeddings,
                loss=loss, losses=losses, counts=counts_all, mmd_rows=mmd_rows)


# ----------------------------------------------------------------------------- a11: classifiers
def class_weights(num_samples, beta: float = 0.999, mult: float = 1e4):
    """models/classifiers/model.py:195-200."""
    import numpy as np
    ns = np.maximum(np.asarray(num_samples, dtype=np.float64), 1e-6)
    eff = 1.0 - np.power(beta, ns * mult)
    w = (1.0 - beta) / eff
    return (w / w.sum() * len(ns)).tolist()


def classifiers_forward(sd, emb: Tensor, labels: Tensor, spec: OracleSpec):
    """9 x Linear(64, C_g) + class-weighted CE; models/classifiers/model.py:74-82, 202-223."""
    x = emb.detach()
    losses, logits, total = {}, {}, 0.0
    for i, key in enumerate(spec.num_classes):
        p = f"classifiers.heads.{key}"
        lg = F.linear(x, sd[f"{p}.layers.0.weight"], sd[f"{p}.layers.0.bias"])
        logits[key] = lg
        ls = F.cross_entropy(lg, labels[..., i], weight=sd[f"{p}.class_weights"])
        losses["clf/" + key] = ls
        total = total + ls
    total = spec.clf_loss_weight * total / len(spec.num_classes)
    losses["clf"] = total
    return dict(logits=logits, loss=total, losses=losses)


# ----------------------------------------------------------------------------- a12: whole step
def scoreperformer_forward(sd, batch: Dict[str, Tensor], spec: OracleSpec, z_prior: Optional[List[Tensor]] = None,
                           training: bool = True, mmd_perm: Optional[List[Optional[Tensor]]] = None):
    """ScorePerformer.forward, default recipe, all dropouts 0; models/scoreperformer/model.py:280-341
    + ScorePerformerMixedLMWrapper.forward (wrappers.py:409-431) + LM loss (wrappers.py:44-59)."""
    score_hidden = encoder_forward(sd, "score_encoder", batch["score"], batch["score_mask"],
                                   list(spec.num_score_tokens), spec.depth_score, spec)
    enc = perf_encoder_forward(sd, batch, spec, z_prior, training, mmd_perm=mmd_perm)

    # MixedLM shift (wrappers.py:409-431)
    seq = batch["perf"][:, :-1]
    seq_masked = batch["masked_perf"][:, 1:]
    labels = batch["labels"][:, 1:]
    context = score_hidden[:, 1:]
    style = enc["embeddings"][:, 1:]
    dmask = batch["perf_mask"][:, :-1]

    x, _ = decoder_embed(sd, seq, seq_masked, context, spec)
    style = style[:, :x.shape[1]]
    dec_hidden, _, _ = transformer_stack(sd, "perf_decoder.model.transformer", x, dmask, spec.depth_dec, True, spec,
                                         style=style)
    logits = tied_head(sd, dec_hidden, spec)

    losses = {}
    for i, key in enumerate(spec.num_tokens):
        if bool(torch.any(labels[..., i] != IGNORE_INDEX)):                         # wrappers.py:56
            losses[key] = F.cross_entropy(logits[key].transpose(1, 2), labels[..., i], ignore_index=IGNORE_INDEX)
    lm_loss = sum(losses.values()) / len(losses)                                      # wrappers.py:59
    loss = lm_loss + enc["loss"]                                                      # model.py:318-321
    losses.update(enc["losses"])

    clf = None
    if spec.num_classes:
        clf_mask = batch["perf_mask"] & ~batch["deadpan_mask"][:, None]             # model.py:325
        clf = classifiers_forward(sd, enc["full_embeddings"][clf_mask], batch["directions"][clf_mask], spec)
        loss = loss + clf["loss"]                                                     # model.py:330-332
        losses.update(clf["losses"])
    return dict(loss=loss, losses=losses, logits=logits, dec_hidden=dec_hidden, score_hidden=score_hidden,
                perf_hidden=enc["hidden_state"], latents=enc["latents"], embeddings=enc["embeddings"],
                counts=enc["counts"], clf_logits=None if clf is None else clf["logits"], lm_loss=lm_loss,
                mmd_rows=enc["mmd_rows"])


# ----------------------------------------------------------------------------- a13: greedy rendering
@torch.no_grad()
def render_greedy(sd, spec: OracleSpec, perf: Tensor, perf_masked: Tensor, score_hidden: Tensor, style: Tensor,
                  mask: Optional[Tensor] = None, use_cache: bool = True) -> Tensor:
    """ScorePerformerMixedLMWrapper.unmask_tokens with filter_kwargs={'k': 1} (greedy), per score.

    models/scoreperformer/wrappers.py:324-407 (+ sampling.py:28-33: top-1 == argmax; PAD/MASK
    logits set to -inf, wrappers.py:364-366).  `perf`/`perf_masked`: [1, T, F]; fields equal to
    MASK(1) in `perf` are filled in position order.  Cached and uncached paths agree.
    """
    out = perf.clone()
    t_total = out.shape[1]
    if mask is None:
        mask = torch.ones(out.shape[:2], dtype=torch.bool, device=out.device)
    unmask = out == 1
    caches = None
    tok_cache = None
    for idx in range(t_total):
        if not bool(unmask[:, idx].any()):
            continue
        fields = torch.where(unmask[0, idx])[0].tolist()
        seq = out[:, :idx + 1][:, :-1]
        seqm = perf_masked[:, :idx + 1][:, 1:]
        ctx = score_hidden[:, 1:idx + 1]
        sty = style[:, 1:idx + 1]
        dm = mask[:, :idx + 1][:, :-1]
        if use_cache and tok_cache is not None:
            n_old = tok_cache.shape[1]
            x_new, te_new = decoder_embed(sd, seq[:, n_old:], seqm[:, n_old:], ctx[:, n_old:], spec)
            assert x_new.shape[1] == 1, "reference cache path advances one position per call"
            x, te = x_new, torch.cat([tok_cache, te_new], dim=1)
            hid, hiddens, kvs = transformer_stack(sd, "perf_decoder.model.transformer", x, dm, spec.depth_dec, True,
                                                  spec, style=sty, caches=caches)
        else:
            x, te = decoder_embed(sd, seq, seqm, ctx, spec)
            hid, hiddens, kvs = transformer_stack(sd, "perf_decoder.model.transformer", x, dm, spec.depth_dec, True,
                                                  spec, style=sty)
        if use_cache:
            tok_cache, caches = te, (hiddens, kvs)
        logits = tied_head(sd, hid[:, idx - 1], spec, keys=fields)
        for f_i, key in zip(fields, logits):
            lg = logits[key].clone()
            lg[:, 0] = -float("inf")
            lg[:, 1] = -float("inf")
            out[:, idx, f_i] = lg.argmax(dim=-1)
    return out


def spec_from_config(cfg) -> OracleSpec:
    """Derive sizes from a resolved model config dict (what Constructor.init receives)."""
    att = cfg["perf_decoder"]["transformer"]["attention"]
    return OracleSpec(
        num_tokens=dict(cfg["num_tokens"]),
        num_score_tokens=dict(cfg.get("num_score_tokens") or cfg["num_tokens"]),
        num_classes=dict((cfg.get("classifiers") or {}).get("num_classes") or {}),
        dim=cfg["dim"],
        emb_dim=cfg["perf_decoder"]["token_embeddings"]["emb_dims"],
        heads=cfg["perf_decoder"]["transformer"]["heads"],
        dim_head=att.get("dim_head", 64),
        depth_score=cfg["score_encoder"]["transformer"]["depth"],
        depth_perf=cfg["perf_encoder"]["transformer"]["depth"],
        depth_dec=cfg["perf_decoder"]["transformer"]["depth"],
        latent_dim=tuple(cfg["perf_encoder"]["latent_dim"]),
        aggregate_mode=tuple(cfg["perf_encoder"]["aggregate_mode"]),
        discrete_ids=tuple(cfg["perf_decoder"]["token_embeddings"]["discrete_ids"]),
        mmd_loss_weight=cfg["perf_encoder"].get("loss_weight", 1.0),
        clf_loss_weight=(cfg.get("classifiers") or {}).get("loss_weight", 1.0),
        deadpan_zero_latent=cfg["perf_encoder"].get("deadpan_zero_latent", False),
    )
