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

import math
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
    """`render_decoder` on the performance decoder of a ScorePerformer model."""
    return render_decoder(model.perf_decoder.model, perf, masked_perf, score_hidden, style, mask=mask, fields=fields,
                          temperature=temperature, top_k=top_k, generator=generator, teacher=teacher, use_graph=use_graph)


@torch.no_grad()
def render_decoder(dec, perf: Tensor, masked_perf: Tensor, score_hidden: Tensor, style: Tensor, mask: Optional[Tensor] = None,
                   fields: Sequence[int] = (3, 5, 10, 11), temperature: float = 1.0, top_k: Optional[int] = 1,
                   generator: Optional[torch.Generator] = None, teacher: Optional[Tensor] = None, use_graph: bool = True,
                   start: int = 0, kv_init: Optional[Sequence[Tensor]] = None, return_kv: bool = False):
    """Fill `fields` of every note >= 1 of `perf` [B, T, F], note by note, for all B scores in lockstep.

    perf / masked_perf: int64 [B, T, F]; score_hidden fp32 [B, T, D]; style fp32 [B, T, S]; mask bool [B, T].
    `top_k=1` is greedy (the parity mode, `filter_kwargs={'k': 1}` in the reference); larger k samples from the top-k
    softmax at `temperature`; `top_k=None` uses the reference default ceil(0.1 * V).
    `teacher` [B, T, F] (optional) is fed as the already-rendered prefix instead of the model's own samples (teacher forcing,
    used by the parity tests to compare every step independently); the returned tensor still holds the model's predictions.
    `use_graph`: capture one note-step in a CUDA graph and replay it (positions are device-side); sampling with a custom
    `generator` runs eagerly.
    Continuation (the streaming `unmask_tokens` call): with `start = s > 0` the notes 0..s are complete in `perf`, `kv_init[l]` bf16
    [B, s, 128] holds the keys | values of decoder positions < s for every layer, and only the positions s..T-2 run (notes s+1..T-1
    are rendered).  `return_kv` also returns the per-layer caches [B, T, 128] (rows < T-1 valid).
    """
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
    assert 0 <= start <= T - 1 and (start == 0 or (kv_init is not None and len(kv_init) == sw.depth))
    for kv, init in zip(kv_caches, kv_init or ()):
        assert init.shape == (B, start, 128)
        kv[:, :start] = init
    out = perf.clone()
    feed = out if teacher is None else teacher
    cat_buf = torch.empty((B, 2 * dec.dim), dtype=BF16, device=dev)
    cat2 = torch.empty((B, 2 * dec.dim), dtype=BF16, device=dev)

    # The position lives on the device: every index below is a device gather / scatter, so ONE captured step replays for
    # all T-1 notes (`use_graph`), instead of ~50 host launches per note.
    pos_t = torch.full((1,), start, dtype=torch.int64, device=dev)           # i
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

    # Everything of a note-step that does not depend on the tokens sampled so far is evaluated for ALL positions before the loop,
    # in large launches: the masked tuple of note i+1 through the embedding and its half of project_multiemb (+ bias), and the
    # context half of project_emb (+ bias).  A step then adds its own half (the previous, complete tuple) as a GEMM with a residual.
    lean = plan is not None and K._os.environ.get("SPB_DECODE_FRONT", "lean") == "lean"
    if lean:
        Dm = dec.dim
        wm_r, wc_l, wc_r = wm16[:, Dm:], wc16[:, :Dm], wc16[:, Dm:]
        # project_emb of the tuple embedding and the left half of project_multiemb are two linear maps in a row: composed once per
        # rendering in fp32 (W = Wm_left Wp, its bias joins the prepared term), so a step needs one GEMM for both
        wm_full = te_mod.project_multiemb.weight.detach().float()
        w_comp16 = K.cast_bf16((wm_full[:, :Dm] @ te_mod.project_emb.weight.detach().float()).contiguous())        # [D, F * 128]
        b_comp = (wm_full[:, :Dm] @ bp.float()).contiguous()                                                          # [D]
        P2 = torch.empty((B, T, Dm), dtype=F32, device=dev)
        C2 = torch.empty((B, T, Dm), dtype=F32, device=dev)
        cb = max(1, min(B, 32768 // max(T, 1)))
        for b0 in range(0, B, cb):
            b1 = min(B, b0 + cb)
            x2, _, _ = K.embed_ln_fwd(masked_perf[b0:b1].reshape(-1, F).contiguous(), table, sizes, ln_w, ln_b)
            x2p = K.gemm(x2, wp16, bias=bp, out_dtype=BF16)
            K.gemm(x2p, wm_r, bias=bm + b_comp, out=P2[b0:b1].view(-1, Dm))
            K.gemm(ctx16[b0:b1].reshape(-1, ctx16.shape[-1]), wc_r, bias=bc, out=C2[b0:b1].view(-1, Dm))
            del x2, x2p
        # ... and the AdaLN terms of every position (one GEMM instead of a tile phase per step), memory permitting
        gb_all = None
        if B * T * plan.gb.shape[1] * 2 <= int(K._os.environ.get("SPB_DECODE_GB_ALL_BYTES", str(8 << 30))):
            gb_all = plan.prepare_adaln(style_f)
        feed_c = feed.contiguous()
        tok_buf = torch.empty((B, F), dtype=feed_c.dtype, device=dev)
        p2_buf = torch.empty((B, Dm), dtype=F32, device=dev)
        c2_buf = torch.empty((B, Dm), dtype=F32, device=dev)
        st_buf = torch.empty((B, style_f.shape[-1]), dtype=F32, device=dev)
        ln_buf = torch.empty((B, Dm), dtype=BF16, device=dev)
        # the rest of the front (te = x1 W^T + p2, emb_norm, project_emb's left half + c2) runs inside the stack kernel when the
        # tuple embedding has the recipe's width; x1 is then a static buffer the embedding kernel fills
        front_in_kernel = gb_all is not None and len(sizes) * 128 == 1536 and Dm == 256 \
            and K._os.environ.get("SPB_DECODE_FRONT_IN_KERNEL", "1") == "1"
        if front_in_kernel:
            x1_buf = torch.empty((B, 1536), dtype=BF16, device=dev)
            plan.set_front(x1_buf, w_comp16, p2_buf, en_w.float().contiguous(), en_b.float().contiguous(), wc_l.t().contiguous(), c2_buf)

    def step_lean():
        # one launch reads what this step needs at the device-side position: tuple i, and the prepared terms / style of note i+1
        if gb_all is None:
            K.gather_at_pos([feed_c, P2, C2, style_f], [tok_buf, p2_buf, c2_buf, st_buf], [0, 1, 1, 1], pos_t)
        else:
            K.gather_at_pos([feed_c, P2, C2], [tok_buf, p2_buf, c2_buf], [0, 1, 1], pos_t)
        if front_in_kernel:
            K.embed_ln_fwd(tok_buf, table, sizes, ln_w, ln_b, out=x1_buf)
            plan.step(None, None, km, pos_t, gb_all=gb_all, use_front=True, barrier_is_zero=True)
        else:
            x1, _, _ = K.embed_ln_fwd(tok_buf, table, sizes, ln_w, ln_b)
            te = K.gemm(x1, w_comp16, residual=p2_buf, out_dtype=F32)
            K.layer_norm_fwd(te, en_w, en_b, out=ln_buf, need_stats=False)
            x = K.gemm(ln_buf, wc_l, residual=c2_buf, out_dtype=F32)
            plan.step(x, st_buf if gb_all is None else None, km, pos_t, gb_all=gb_all)
        # tied head for the masked fields only (wrappers.py:364-380); the stack kernel leaves a bf16 copy of its output
        e_raw = K.gemm(plan.out16, whead16, trans_b=True, out_dtype=BF16)
        e, _, _ = K.layer_norm_fwd(e_raw, hn_w, hn_b, out_dtype=BF16, need_stats=False)
        if front_in_kernel:
            # the sampling kernel's last CTA ends the step: position += 1, grid-barrier word of the stack kernel back to zero
            K.sample_fields(e, table16, fields, [offs[f] for f in fields], [sizes[f] for f in fields],
                            [top_k if top_k is not None else -(-sizes[f] // 10) for f in fields], out, pos_t,
                            temperature=temperature, seed=seed, advance=plan.advance)
        else:
            K.sample_fields(e, table16, fields, [offs[f] for f in fields], [sizes[f] for f in fields],
                            [top_k if top_k is not None else -(-sizes[f] // 10) for f in fields], out, pos_t,
                            temperature=temperature, seed=seed)
            pos_t.add_(1)

    def step():
        if lean and fused_sampling and (feed_c is out or teacher is not None):
            return step_lean()
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

    n_steps = T - 1 - start
    graph_ok = use_graph and generator is None and n_steps > 4
    if not graph_ok:
        for _ in range(n_steps):
            step()
        return (out, kv_caches) if return_kv else out
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
    return (out, kv_caches) if return_kv else out


# ----------------------------------------------------------------------------- reference-signature entry points
# `unmask_tokens` / `generate` of the LM wrappers (models/scoreperformer/wrappers.py:200-307, 324-407 of the reference) keep their
# argument lists and their cache contract, so inference/generators.py works unchanged -- but they are adapters over this file:
# a window goes through `render_decoder`'s device-resident loop whenever the request is expressible there (any batch size,
# top-k / greedy sampling, no per-key bans) -- from note 1 for a fresh window, or from the end of caller-held caches whose length
# is the known prefix (the streaming calls of inference/generators.py); everything else runs the general stepper below, which
# advances one note per iteration through the cached stack step and samples on the host side of the logits.
def _sampling_plan(filter_logits_fn, filter_kwargs, sizes: Sequence[int], fields: Sequence[int]):
    """k per field if the filter is the reference's `top_k` (sampling.py:28-33), else None (host sampling)."""
    from .modules import sampling
    if filter_logits_fn is not sampling.top_k:
        return None
    kw = dict(filter_kwargs or {})
    if set(kw) - {"k", "thres"}:
        return None
    ks = [int(kw["k"]) if kw.get("k") is not None else math.ceil((1 - kw.get("thres", 0.9)) * sizes[f]) for f in fields]
    return ks if all(1 <= k <= 32 for k in ks) else None


def _draw(logits: Dict[str, Tensor], banned: Sequence[int], filter_key_ids, filter_logits_fn, filter_kwargs, temperature: float) -> Tensor:
    """One token per field from fp32 logits [B, V]: bans, the caller's filter, softmax at the temperature, a draw."""
    from .modules.sampling import filter_logits_and_sample
    cols = []
    for key, lg in logits.items():
        lg = lg.clone()
        lg[:, list(banned)] = -float("inf")
        extra = (filter_key_ids or {}).get(key)
        if extra is not None:
            lg[:, extra] = -float("inf")
        cols.append(filter_logits_and_sample(lg, filter_logits_fn, filter_kwargs=filter_kwargs, temperature=temperature))
    return torch.cat(cols, dim=-1)


def _masked_columns(tokens: Tensor, mask_token_id: int):
    """(note indices that contain a MASK, bool [F] of the masked fields) -- the same fields at every note of the window, as the
    collator produces them; raises otherwise."""
    hit = tokens == mask_token_id
    notes = torch.nonzero(hit.any(dim=2).any(dim=0)).flatten()
    cols = hit.any(dim=1).any(dim=0)
    return notes, cols, hit


@torch.inference_mode()
def unmask_mixlm(wrapper, tokens: Tensor, tokens_masked: Tensor, temperature: float, filter_logits_fn, filter_kwargs, filter_key_ids,
                 caches, return_caches: bool, **kwargs):
    """MixedLM rendering with the reference's signature (wrappers.py:324-407): fill the MASKed fields of `tokens` note by note;
    position i of the decoder sees the full tuple of note i and the masked tuple of note i+1."""
    dec = wrapper.model
    was_training = dec.training
    dec.eval()
    squeeze = tokens.dim() == 2
    if squeeze:
        tokens, tokens_masked = tokens[None], tokens_masked[None]
    out = tokens.clone()
    mask = kwargs.pop("mask", None)
    notes, cols, hit = _masked_columns(out, wrapper.mask_token_id)
    fields = torch.nonzero(cols).flatten().tolist()
    sizes = dec.token_emb.field_sizes
    ks = _sampling_plan(filter_logits_fn, filter_kwargs, sizes, fields)
    context, style = kwargs.get("context"), kwargs.get("style_embeddings")
    T = out.shape[1]
    contiguous_tail = notes.numel() > 0 and int(notes[0]) >= 1 and torch.equal(notes, torch.arange(int(notes[0]), T, device=notes.device))
    uniform = bool(hit[:, notes][..., cols].all()) if notes.numel() else False
    expressible = (ks is not None and not filter_key_ids and contiguous_tail and uniform and context is not None and style is not None
                   and len(set(ks)) == 1 and set(kwargs) <= {"context", "style_embeddings"})
    fast = expressible and caches is None and not return_caches and int(notes[0]) == 1
    kv_init = _streamable(dec, caches, int(notes[0]) - 1, context, style, T) if expressible and not fast else None
    if fast:
        res = render_decoder(dec, out, tokens_masked, context, style, mask=mask, fields=fields, temperature=temperature, top_k=ks[0])
    elif kv_init is not None:
        # the streaming call of inference/generators.py: a few new notes behind a prefix whose keys / values the caller holds.  Same
        # device-resident note-step as a whole-window rendering (persistent stack kernel, fused heads + sampling), started at the
        # first new note; eager launches, a chord is too short to pay for a graph capture
        res, kvs = render_decoder(dec, out, tokens_masked, context, style, mask=mask, fields=fields, temperature=temperature,
                                  top_k=ks[0], use_graph=False, start=int(notes[0]) - 1, kv_init=kv_init, return_kv=True)
        caches = _kv_only_caches(dec, kvs, T - 1)
    else:
        res, caches = _unmask_stepwise(wrapper, out, tokens_masked, mask, notes, hit, temperature, filter_logits_fn, filter_kwargs,
                                       filter_key_ids, caches, kwargs)
    dec.train(was_training)
    res = res[0] if squeeze else res
    return (res, caches) if return_caches else res


def _streamable(dec, caches, start: int, context: Tensor, style: Tensor, T: int):
    """Per-layer [B, start, 128] keys | values if the request can continue on the device-resident note-step: nothing cached and the
    first note is the one to render, or caches whose length is exactly the known prefix (the check inference/generators.py:222-226
    makes as well).  None sends the request to the general stepper.  SPB_STREAM=legacy switches this path off."""
    if K._os.environ.get("SPB_STREAM", "fused") != "fused" or context.shape[1] != T or style.shape[1] != T or start < 0:
        return None
    if caches is None:
        return [] if start == 0 else None
    att = caches.transformer.attention if caches.transformer is not None else None
    if not att or len(att) != dec.transformer.depth or caches.token_emb.shape[1] != start or start == 0:
        return None
    if any(a.keys is None or a.values is None or a.keys.shape[-2] != start for a in att):
        return None
    return [torch.cat([a.keys, a.values], dim=-1).to(BF16).contiguous() for a in att]


def _kv_only_caches(dec, kvs: Sequence[Tensor], length: int):
    """Caches in the reference's layout (TupleTransformerCaches / TransformerIntermediates / AttentionIntermediates, every tensor
    [B, length, .]) after a device-resident rendering.  Keys and values are views of the working buffers.  `token_emb` and `hiddens`
    are ZERO placeholders of the right shapes: the cached step of this implementation needs the keys and values of the prefix and
    the last position only (the reference recomputes emb_norm / project_emb over the whole prefix from `token_emb` at every step,
    transformer.py:171-185, and carries the per-layer hiddens without reading them)."""
    from .models.scoreperformer.transformer import TupleTransformerCaches
    from .modules.transformer.attend import AttentionIntermediates
    from .modules.transformer.transformer import TransformerIntermediates
    B = kvs[0].shape[0]
    z = torch.zeros((B, length, dec.dim), dtype=F32, device=kvs[0].device)
    att = [AttentionIntermediates(keys=kv[:, :length, :64], values=kv[:, :length, 64:]) for kv in kvs]
    return TupleTransformerCaches(token_emb=z, transformer=TransformerIntermediates(hiddens=[z] * (len(kvs) + 1), attention=att))


def _unmask_stepwise(wrapper, out, tokens_masked, mask, notes, hit, temperature, filter_logits_fn, filter_kwargs, filter_key_ids, caches, kwargs):
    """General stepper: one cached forward per note (the wrapper's own `forward` slices context / style / masks the way training
    does), logits of the note's masked fields from the tied head, host-side filter + draw."""
    names = list(wrapper.model.lm_head.embs.keys())
    if mask is None:
        mask = torch.ones(out.shape[:2], dtype=torch.bool, device=out.device)
    for idx in notes.tolist():
        keys = torch.nonzero(hit[0, idx]).flatten().tolist()
        step = wrapper(out[:, :idx + 1], seq_masked=tokens_masked[:, :idx + 1], mask=mask[:, :idx + 1], return_embeddings=True,
                       return_caches=True, caches=caches, **kwargs)
        caches = step.caches
        logits = wrapper.model.lm_head(step.hidden_state[:, idx - 1], keys=keys)
        out[:, idx, keys] = _draw(logits, (wrapper.pad_token_id, wrapper.mask_token_id), filter_key_ids, filter_logits_fn, filter_kwargs,
                                  temperature)
    return out, caches


@torch.inference_mode()
def unmask_mlm(wrapper, tokens: Tensor, single_run: bool, temperature: float, filter_logits_fn, filter_kwargs, filter_key_ids, **kwargs):
    """Masked-LM unmasking (wrappers.py:131-198 of the reference): every MASK at once from one forward (`single_run`, arg-max),
    or note by note re-running the (uncached, bidirectional) model."""
    dec = wrapper.model
    was_training = dec.training
    dec.eval()
    squeeze = tokens.dim() == 2
    out = (tokens[None] if squeeze else tokens).clone()
    mask = kwargs.pop("mask", None)
    if mask is None:
        mask = torch.ones(out.shape[:2], dtype=torch.bool, device=out.device)
    notes, _, hit = _masked_columns(out, wrapper.mask_token_id)
    if single_run:
        logits = dec(out, mask=mask, **kwargs).logits
        best = torch.stack([lg.argmax(dim=-1) for lg in logits.values()], dim=-1)
        out = torch.where(hit, best, out)
    else:
        for idx in notes.tolist():
            keys = torch.nonzero(hit[0, idx]).flatten().tolist()
            step = wrapper(out[:, :idx + 1], mask=mask[:, :idx + 1], return_embeddings=True, **kwargs)
            logits = dec.lm_head(step.hidden_state[:, idx - 1], keys=keys)
            out[:, idx, keys] = _draw(logits, range(wrapper.num_special_tokens), filter_key_ids, filter_logits_fn, filter_kwargs, temperature)
    dec.train(was_training)
    return out[0] if squeeze else out


@torch.inference_mode()
def generate_ar(wrapper, start_tokens: Tensor, seq_len: int, max_bar, temperature: float, filter_logits_fn, filter_kwargs, caches,
                return_caches: bool, tokenizer, fix_errors: bool, **kwargs):
    """Causal-LM continuation (wrappers.py:200-307 of the reference): append one sampled tuple per step through the cached stack
    step until `seq_len`, EOS, or a bar beyond `max_bar`.  With a tokenizer and `fix_errors` the bar never decreases and the
    tempo / time signature are held inside a bar."""
    dec = wrapper.model
    was_training = dec.training
    dec.eval()
    squeeze = start_tokens.dim() == 2
    out = start_tokens[None] if squeeze else start_tokens
    t0 = out.shape[1]
    mask = kwargs.pop("mask", None)
    if mask is None:
        mask = torch.ones(out.shape[:2], dtype=torch.bool, device=out.device)
    col = getattr(tokenizer, "vocab_types_idx", None) if fix_errors else None
    for _ in range(t0, seq_len + 1):
        window, wmask = out[:, -wrapper.max_seq_len:], mask[:, -wrapper.max_seq_len:]
        step = wrapper(window, mask=wmask, caches=caches, return_embeddings=True, return_caches=True, **kwargs)
        caches = step.caches
        logits = dec.lm_head(step.hidden_state[:, -1])
        last = out[:, -1]
        new, bar_tok = [], None
        for key, lg in logits.items():
            if col is not None and key == "Bar":
                lg = lg.clone()
                lg[:, 4:int(last[0, col["Bar"]])] = -float("inf")           # bars only move forward
            held = col is not None and (key == "TimeSig" or (key == "Tempo" and bar_tok is not None and bool((bar_tok == last[:, col["Bar"]]).all())))
            tok = last[:, col[key]][:, None] if held else _draw({key: lg}, (0, 1), None, filter_logits_fn, filter_kwargs, temperature)
            if key == "Bar":
                bar_tok = tok[:, 0]
            new.append(tok)
        out = torch.cat([out, torch.cat(new, dim=-1)[:, None]], dim=1)
        mask = torch.nn.functional.pad(mask, (0, 1), value=True)
        if wrapper.eos_token_id is not None:
            if bool((out[:, -1, 0] == wrapper.eos_token_id).any()):
                out[:, -1, 1:] = wrapper.pad_token_id
                break
        elif max_bar is not None and bool((out[:, -1, 0] > max_bar).any()):
            out = out[:, :-1]
            break
    out = out[:, t0:]
    dec.train(was_training)
    out = out[0] if squeeze else out
    return (out, caches) if return_caches else out
