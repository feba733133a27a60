"""KV-cached incremental decoding.

Two entry points over the same kernels:

* `cached_stack_step` -- the cache path of `Transformer.forward` (modules/transformer/transformer.py:161-186,219-221) with
  the reference's cache contract (`TransformerIntermediates`: per attention layer the input hiddens `[B,t,D]` and
  `AttentionIntermediates(keys[B,t,64], values[B,t,64])`), used by `unmask_tokens` / `generate` and therefore by an
  unchanged `inference/generators.py`.
* `render_batch` -- the batched, device-resident form of `ScorePerformerMixedLMWrapper.unmask_tokens`
  (models/scoreperformer/wrappers.py:324-407): all scores advance in lockstep, KV caches are preallocated
  `[depth, B, T, 128]` buffers written in place (no per-step `torch.cat`), no host sync inside the loop.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import torch
from torch import Tensor

from . import fused, kernels as K

BF16, F32 = torch.bfloat16, torch.float32


class _StackWeights:
    """bf16 copies of one stack's weights, prepared once per decode session."""

    def __init__(self, tr):
        tr._check_fused()
        self.depth, self.H, self.D = tr.depth, tr.heads, tr.dim
        self.ada = tr.ada_norm
        self.layers = []
        norms_w, norms_b = [], []
        for l in range(tr.depth):
            na, attn, _ = tr.layers[2 * l]
            nf, ff, _ = tr.layers[2 * l + 1]
            for n in (na[0], nf[0]):
                w, b = tr._norm_params(n)
                norms_w.append(w.detach())
                norms_b.append(b.detach())
            self.layers.append(dict(
                wqkv=K.cast_bf16(torch.cat([attn.to_q.weight, attn.to_k.weight, attn.to_v.weight], 0).detach().contiguous()),
                wo=K.cast_bf16(attn.to_out.weight.detach().contiguous()),
                ls=attn.logslopes().detach().reshape(-1).float().contiguous(),
                w1=K.cast_bf16(ff.ff[0].proj.weight.detach().contiguous()), b1=ff.ff[0].proj.bias.detach().float().contiguous(),
                w2=K.cast_bf16(ff.ff[3].weight.detach().contiguous())))
        w, b = tr._norm_params(tr.final_norm)
        norms_w.append(w.detach())
        norms_b.append(b.detach())
        if self.ada:
            self.w_ada = K.cast_bf16(torch.cat(norms_w, 0).contiguous())
            self.b_ada = fused.ada_bias_minus_one(norms_b, self.D).contiguous()      # gb = (gamma - 1 | beta), see rowops.cu
        else:
            self.norm_w, self.norm_b = [w.float().contiguous() for w in norms_w], [b.float().contiguous() for b in norms_b]

    def norm(self, i: int, x: Tensor, gb_all: Optional[Tensor], out_dtype=BF16) -> Tensor:
        if self.ada:
            gb = gb_all[:, i * 2 * self.D:(i + 1) * 2 * self.D]
            return K.layer_norm_fwd(x, None, None, gb, out_dtype=out_dtype, need_stats=False)[0]
        return K.layer_norm_fwd(x, self.norm_w[i], self.norm_b[i], out_dtype=out_dtype, need_stats=False)[0]


def _stack_step(sw: _StackWeights, x_last: Tensor, style_last: Optional[Tensor], kv_caches: Sequence[Tensor], key_mask: Optional[Tensor],
                pos: int, hid_out: Optional[List[Tensor]] = None, pos_dev: Optional[Tensor] = None) -> Tensor:
    """One new position through the stack.  x_last fp32 [B, D]; kv_caches[l] bf16 [B, cap, 128] already holding rows < pos;
    row `pos` is written here (by the attention kernel itself).  With `pos_dev` (device int64 [1]) the position is only known on
    the device -- `pos` is ignored and every launch is CUDA-graph replayable.  Returns the final-norm output fp32 [B, D]."""
    gb_all = None
    if sw.ada:
        gb_all = K.gemm(K.cast_bf16(style_last.float().contiguous()), sw.w_ada, bias=sw.b_ada, out_dtype=BF16)
    if key_mask is None:
        rowmask = None
    elif pos_dev is None:
        rowmask = key_mask[:, pos].contiguous()
    else:
        rowmask = key_mask.index_select(1, pos_dev).reshape(-1).contiguous()
    cur = x_last
    for l, w in enumerate(sw.layers):
        if hid_out is not None:
            hid_out.append(cur)
        xn = sw.norm(2 * l, cur, gb_all)
        qkv = K.gemm(xn, w["wqkv"], out_dtype=BF16)
        cap = kv_caches[l].shape[1]
        o = K.attention_decode(qkv, kv_caches[l], key_mask, w["ls"], sw.H, cap if pos_dev is not None else pos + 1, pos,
                               pos_dev=pos_dev, append_kv=True)
        cur = K.gemm(o, w["wo"], residual=cur, rowmask=rowmask, out_dtype=F32)
        xn = sw.norm(2 * l + 1, cur, gb_all)
        u = K.gemm(xn, w["w1"], bias=w["b1"], out_dtype=BF16)
        h = K.glu_fwd(u, 0.0, 0)
        cur = K.gemm(h, w["w2"], residual=cur, out_dtype=F32)
    return sw.norm(2 * sw.depth, cur, gb_all, out_dtype=F32)


@torch.no_grad()
def cached_stack_step(tr, x: Tensor, mask: Optional[Tensor], style: Optional[Tensor], cache, return_hiddens: bool):
    """Reference cache contract: only the last position of `x` is new; returns the full-length output (and caches)."""
    from .modules.transformer.attend import AttentionIntermediates
    from .modules.transformer.transformer import TransformerIntermediates
    b, t, d = x.shape
    sw = _StackWeights(tr)
    assert len(cache.attention) == tr.depth and len(cache.hiddens) == tr.depth + 1
    kvs = []
    for inter in cache.attention:                         # [B, t-1, 64] each -> working buffer [B, t, 128]
        kv = torch.empty((b, t, 128), dtype=BF16, device=x.device)
        kv[:, :t - 1, :64] = inter.keys
        kv[:, :t - 1, 64:] = inter.values
        kvs.append(kv)
    hid_new: List[Tensor] = []
    km = None if mask is None else mask.contiguous()
    out_last = _stack_step(sw, x[:, -1].float().contiguous(), None if style is None else style[:, -1], kvs, km, t - 1, hid_new)
    out = torch.cat([cache.hiddens[-1].float(), out_last[:, None]], dim=1)
    if not return_hiddens:
        return out
    hiddens = [torch.cat([c.float(), h[:, None]], dim=1) for c, h in zip(cache.hiddens[:-1], hid_new)] + [out]
    att = [AttentionIntermediates(keys=kv[..., :64], values=kv[..., 64:]) for kv in kvs]
    return out, TransformerIntermediates(hiddens=hiddens, attention=att)


@torch.no_grad()
def render_batch(model, perf: Tensor, masked_perf: Tensor, score_hidden: Tensor, style: Tensor, mask: Optional[Tensor] = None,
                 fields: Sequence[int] = (3, 5, 10, 11), temperature: float = 1.0, top_k: Optional[int] = 1,
                 generator: Optional[torch.Generator] = None, teacher: Optional[Tensor] = None, use_graph: bool = True) -> Tensor:
    """Fill `fields` of every note >= 1 of `perf` [B, T, F], note by note, for all B scores in lockstep.

    perf / masked_perf: int64 [B, T, F]; score_hidden fp32 [B, T, D]; style fp32 [B, T, S]; mask bool [B, T].
    `top_k=1` is greedy (the parity mode, `filter_kwargs={'k': 1}` in the reference); larger k samples from the top-k
    softmax at `temperature`; `top_k=None` uses the reference default ceil(0.1 * V).
    `teacher` [B, T, F] (optional) is fed as the already-rendered prefix instead of the model's own samples (teacher forcing,
    used by the parity tests to compare every step independently); the returned tensor still holds the model's predictions.
    `use_graph`: capture one note-step in a CUDA graph and replay it (positions are device-side); sampling with a custom
    `generator` runs eagerly.
    """
    dec = model.perf_decoder.model
    te_mod, head = dec.token_emb, dec.lm_head
    B, T, F = perf.shape
    dev = perf.device
    sw = _StackWeights(dec.transformer)
    table = te_mod.table().detach().contiguous()
    table16 = K.cast_bf16(table)
    sizes = te_mod.field_sizes
    ln_w, ln_b = te_mod.norm.weight.detach(), te_mod.norm.bias.detach()
    wp16 = K.cast_bf16(te_mod.project_emb.weight.detach().contiguous())
    bp = te_mod.project_emb.bias.detach()
    wm16 = K.cast_bf16(te_mod.project_multiemb.weight.detach().contiguous())
    bm = te_mod.project_multiemb.bias.detach()
    en_w, en_b = dec.emb_norm.weight.detach(), dec.emb_norm.bias.detach()
    wc16 = K.cast_bf16(dec.project_emb.weight.detach().contiguous())
    bc = dec.project_emb.bias.detach()
    whead16 = K.cast_bf16(head.proj_weight_kn().detach().contiguous())
    hn_w, hn_b = head.norm.weight.detach(), head.norm.bias.detach()
    emb = head.split_dims[0]
    offs = [0]
    for v in sizes[:-1]:
        offs.append(offs[-1] + v)
    ctx16 = K.cast_bf16(score_hidden.float().contiguous().view(B * T, -1)).view(B, T, -1)
    km = None if mask is None else mask.contiguous()
    kv_caches = [torch.zeros((B, T, 128), dtype=BF16, device=dev) for _ in range(sw.depth)]
    out = perf.clone()
    feed = out if teacher is None else teacher
    cat_buf = torch.empty((B, 2 * dec.dim), dtype=BF16, device=dev)
    cat2 = torch.empty((B, 2 * dec.dim), dtype=BF16, device=dev)

    # The position lives on the device: every index below is a device gather / scatter, so ONE captured step replays for
    # all T-1 notes (`use_graph`), instead of ~50 host launches per note.
    pos_t = torch.zeros(1, dtype=torch.int64, device=dev)                    # i
    field_idx = torch.tensor(list(fields), dtype=torch.int64, device=dev)
    neg_inf = -float("inf")

    # the whole decoder stack of a note-step is ONE persistent kernel when the stack has the recipe's shape (csrc/decode_stack.cu);
    # SPB_DECODE=legacy keeps the launch-per-operator path (also the fallback for other shapes)
    plan = None
    style_f = style.float().contiguous()
    if K.decode_stack_ok(sw.depth, sw.D, sw.H, sw.layers[0]["w2"].shape[1], sw.ada, style_f.shape[-1], T):
        plan = K.DecodeStackPlan(sw.layers, sw.w_ada, sw.b_ada, kv_caches, B, style_f.shape[-1])

    ks = [top_k if top_k is not None else -(-sizes[f] // 10) for f in fields]
    fused_sampling = (generator is None and emb == 128 and len(fields) <= 8 and all(sizes[f] <= 256 for f in fields)
                      and all(1 <= k <= 32 for k in ks) and K._os.environ.get("SPB_DECODE", "fused") == "fused")
    seed = K.seed_from_torch() if any(k > 1 for k in ks) else 0

    def step():
        nxt = pos_t + 1
        # decoder position i: full tuple of note i, masked tuple / context / style of note i+1 (wrappers.py:409-431)
        x1, _, _ = K.embed_ln_fwd(feed.index_select(1, pos_t).reshape(B, F).contiguous(), table, sizes, ln_w, ln_b)
        x2, _, _ = K.embed_ln_fwd(masked_perf.index_select(1, nxt).reshape(B, F).contiguous(), table, sizes, ln_w, ln_b)
        K.gemm(x1, wp16, bias=bp, out=cat_buf[:, :dec.dim])
        K.gemm(x2, wp16, bias=bp, out=cat_buf[:, dec.dim:])
        te = K.gemm(cat_buf, wm16, bias=bm, out_dtype=F32)
        K.layer_norm_fwd(te, en_w, en_b, out=cat2[:, :dec.dim], need_stats=False)
        cat2[:, dec.dim:].copy_(ctx16.index_select(1, nxt).reshape(B, -1))
        x = K.gemm(cat2, wc16, bias=bc, out_dtype=F32)
        if plan is not None:
            hid = plan.step(x, style_f.index_select(1, nxt).reshape(B, -1), km, pos_t)
        else:
            hid = _stack_step(sw, x, style.index_select(1, nxt).reshape(B, -1), kv_caches, km, 0, pos_dev=pos_t)
        # tied head for the masked fields only (wrappers.py:364-380)
        e_raw = K.gemm(K.cast_bf16(hid), whead16, trans_b=True, out_dtype=BF16)
        e, _, _ = K.layer_norm_fwd(e_raw, hn_w, hn_b, out_dtype=BF16, need_stats=False)
        if fused_sampling:
            # heads of the rendered fields, PAD / MASK ban, top-k filter, draw and token write: one launch
            K.sample_fields(e, table16, fields, [offs[f] for f in fields], [sizes[f] for f in fields],
                            [top_k if top_k is not None else -(-sizes[f] // 10) for f in fields], out, pos_t,
                            temperature=temperature, seed=seed)
        else:
            toks = []
            for f in fields:
                lg = K.gemm(e[:, f * emb:(f + 1) * emb], table16[offs[f]:offs[f] + sizes[f]], out_dtype=F32)
                lg[:, :2] = neg_inf                                             # PAD / MASK are never emitted
                k = top_k if top_k is not None else -(-sizes[f] // 10)
                if k == 1:
                    tok = lg.argmax(dim=-1)
                else:
                    val, ind = torch.topk(lg, k)
                    probs = torch.softmax(val / temperature, dim=-1)
                    tok = ind.gather(1, torch.multinomial(probs, 1, generator=generator)).squeeze(1)
                toks.append(tok)
            # out[:, i+1, fields] = toks
            out_rows = out.view(B, T * F)
            dst = nxt * F + field_idx                                           # [n_fields] flat column indices
            out_rows.index_copy_(1, dst, torch.stack(toks, dim=1))
        pos_t.add_(1)

    n_steps = T - 1
    graph_ok = use_graph and generator is None and n_steps > 4
    if not graph_ok:
        for _ in range(n_steps):
            step()
        return out
    # two eager steps warm every lazily-initialised path, then one step is captured and replayed for the rest
    step()
    step()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    before = K.LAUNCHES
    with torch.cuda.graph(graph):
        step()                                   # recorded, not executed
    per_step = K.LAUNCHES - before
    K.LAUNCHES = before                          # the recording itself launched nothing ...
    for _ in range(n_steps - 2):
        graph.replay()
    K.LAUNCHES += per_step * (n_steps - 2)       # ... each replay launches every recorded kernel
    return out
