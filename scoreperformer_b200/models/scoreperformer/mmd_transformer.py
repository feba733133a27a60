"""TupleTransformer with hierarchical MMD-VAE latent heads.

Reference: scoreperformer/models/scoreperformer/mmd_transformer.py:18-542.  The per-level pooling / projection /
broadcast-back and the MMD loss run as fused kernels (fused.LatentLevelsFn, fused.MMDFn); this file keeps the
reference's constructor, output dataclass, loss keys and the pure helper methods used by the generator.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Union

import torch
import torch.nn as nn
from torch import Tensor

from ... import fused
from ...modules.transformer import TransformerConfig
from ...utils import ExplicitEnum, SideBranch
from .embeddings import TupleTokenEmbeddingsConfig
from .transformer import TupleTransformer, TupleTransformerConfig, TupleTransformerOutput


class EmbeddingAggregateModes(ExplicitEnum):
    SAME = "same"
    MEAN = "mean"
    BEAT_MEAN = "beat_mean"
    BAR_MEAN = "bar_mean"
    ONSET_MEAN = "onset_mean"
    ISOLATED_BAR_MEAN = "isolated_bar_mean"


SEGMENT_MODES = ("isolated_bar_mean", "bar_mean", "beat_mean", "onset_mean")


@dataclass
class MMDTupleTransformerOutput(TupleTransformerOutput):
    latents: Optional[Union[Tensor, List[Tensor]]] = None
    embeddings: Optional[Tensor] = None
    full_embeddings: Optional[Tensor] = None
    dropout_mask: Optional[Tensor] = None
    loss: Optional[Tensor] = None
    losses: Optional[Dict[str, Tensor]] = None


@dataclass
class MMDTupleTransformerConfig(TupleTransformerConfig):
    latent_dim: Union[int, List[int]] = 64
    aggregate_mode: Union[str, List[str]] = EmbeddingAggregateModes.MEAN
    hierarchical: bool = False
    hierarchical_with_context: bool = True
    latent_dropout: Union[float, List[float]] = 0.
    inclusive_latent_dropout: bool = True
    deadpan_zero_latent: bool = False
    loss_weight: float = 1.0


class MMDVAE(nn.Module):
    def __init__(self, input_dim, latent_dim):
        super().__init__()
        self.latent_dim = latent_dim
        self.linear = nn.Linear(input_dim, latent_dim)

    def forward(self, inputs: Tensor):
        return fused.linear(inputs, self.linear.weight, self.linear.bias, out_fp32=True)


class MMDLoss(nn.Module):
    """MMD between N(0, I) samples and the valid latents (mmd_transformer.py:505-534), one fused pairwise-RBF kernel."""

    def __init__(self, num_samples: int = 256, max_num_latents: int = 4096):
        super().__init__()
        self.num_samples = num_samples
        self.max_num_latents = max_num_latents

    def forward(self, latents: Tensor, mask: Optional[Tensor] = None, z: Optional[Tensor] = None, rows: Optional[Tensor] = None):
        """latents [..., d]; mask [...] bool (True = valid).  `z` injects the prior sample and `rows` ([n, 2] = (sample,
        segment) pairs) the `randperm` subsample of mmd_transformer.py:515-517 (parity runs)."""
        d = latents.shape[-1]
        y = latents.reshape(-1, d)
        w = torch.ones(y.shape[0], dtype=torch.bool, device=y.device) if mask is None else mask.reshape(-1)
        if rows is not None:
            idx = rows[:, 0].to(y.device) * latents.shape[-2] + rows[:, 1].to(y.device)
            y, w = y.index_select(0, idx), w.index_select(0, idx)
        elif y.shape[0] > self.max_num_latents:
            # device-side replacement of `latents[mask][randperm(n)[:max]]`: the `max` smallest random keys among the valid
            # rows (all valid rows when n <= max) -- same distribution, no host sync, static shapes.
            keys = torch.rand(y.shape[0], device=y.device).masked_fill_(~w, 2.0)
            idx = torch.topk(keys, self.max_num_latents, largest=False, sorted=False).indices
            y, w = y.index_select(0, idx), w.index_select(0, idx)
        if z is None:
            z = torch.randn(self.num_samples, d, device=y.device, dtype=torch.float32)
        return fused.MMDFn.apply(y.float(), w, z.float())

    @staticmethod
    def gaussian_kernel(x, y):
        num = (x.unsqueeze(1) - y.unsqueeze(0)).pow(2).mean(2) / x.size(-1)
        return torch.exp(-num)

    @staticmethod
    def compute_mmd(x, y):
        return MMDLoss.gaussian_kernel(x, x).mean() + MMDLoss.gaussian_kernel(y, y).mean() - 2 * MMDLoss.gaussian_kernel(x, y).mean()


class MMDTupleTransformer(TupleTransformer):
    def __init__(self, num_tokens: Dict[str, int], dim: int = 512, max_seq_len: int = 1024, transformer=None, token_embeddings=None,
                 use_abs_pos_emb: bool = True, emb_norm: bool = False, emb_dropout: float = 0.0, context_emb_dim: Optional[int] = None,
                 context_emb_mode: str = "attention", style_emb_dim: Optional[int] = None, style_emb_mode: str = "cat", lm_head=None,
                 regression_head=None, latent_dim: Union[int, List[int]] = 64, aggregate_mode=EmbeddingAggregateModes.MEAN,
                 hierarchical: bool = False, hierarchical_with_context: bool = True, latent_dropout: Union[float, List[float]] = 0.,
                 inclusive_latent_dropout: bool = True, deadpan_zero_latent: bool = False, loss_weight: float = 1.0):
        if transformer is None:
            transformer = TransformerConfig(_target_="default")
        if token_embeddings is None:
            token_embeddings = TupleTokenEmbeddingsConfig()
        super().__init__(num_tokens=num_tokens, dim=dim, max_seq_len=max_seq_len, transformer=transformer,
                         token_embeddings=token_embeddings, use_abs_pos_emb=use_abs_pos_emb, emb_norm=emb_norm, emb_dropout=emb_dropout,
                         context_emb_dim=context_emb_dim, context_emb_mode=context_emb_mode, style_emb_dim=style_emb_dim,
                         style_emb_mode=style_emb_mode, lm_head=lm_head, regression_head=regression_head)

        if not isinstance(latent_dim, int):
            latent_dim = list(latent_dim)
            aggregate_mode = [aggregate_mode] * len(latent_dim) if isinstance(aggregate_mode, str) else list(aggregate_mode)
        if isinstance(aggregate_mode, str):
            assert EmbeddingAggregateModes.has_value(aggregate_mode), \
                f"`{aggregate_mode}` is not a valid aggregate_mode`, available modes: {EmbeddingAggregateModes.list()}"
        else:
            aggregate_mode = list(aggregate_mode)
            latent_dim = [latent_dim] * len(aggregate_mode) if isinstance(latent_dim, int) else latent_dim
            for mode in aggregate_mode:
                assert EmbeddingAggregateModes.has_value(mode), \
                    f"`{mode}` is not a valid aggregate_mode`, available modes: {EmbeddingAggregateModes.list()}"
        assert not hierarchical or isinstance(aggregate_mode, list), "`hierarchical` mode can only be used with multiple VAE heads"
        self.hierarchical = hierarchical
        self.hierarchical_with_context = hierarchical_with_context
        if not isinstance(latent_dim, int):
            latent_dropout = [latent_dropout] * len(latent_dim) if isinstance(latent_dropout, (int, float)) else list(latent_dropout)

        self.aggregate_mode = aggregate_mode
        self.latent_dim = latent_dim
        self.latent_dropout = latent_dropout
        self.inclusive_latent_dropout = inclusive_latent_dropout
        self.deadpan_zero_latent = deadpan_zero_latent

        if isinstance(latent_dim, int):
            self.vae_head = MMDVAE(input_dim=dim, latent_dim=latent_dim)
            self.embedding_dim = latent_dim
        else:
            self.vae_head = nn.ModuleDict()
            input_dim = dim
            for mode, z in zip(aggregate_mode, latent_dim):
                self.vae_head[mode] = MMDVAE(input_dim=input_dim, latent_dim=z)
                if self.hierarchical:
                    input_dim = input_dim + z if self.hierarchical_with_context else z
            self.embedding_dim = sum(latent_dim)

        self.criterion = MMDLoss()
        self.loss_weight = loss_weight
        self.pad_token_id, self.mask_token_id, self.sos_token_id, self.eos_token_id = 0, 1, 2, 3
        self._mask_bars = False
        # When False the segment tables are sized T + 4 (static, no host sync); when True one `.max()` sync per forward
        # trims `latents` to the reference's exact [B, max_id + 1, z] shape.
        self.exact_latent_shapes = True
        self.slot_capacity = None          # rows of every segment table in sync-free mode (None: T + 4)
        self.segment_overflow = None       # device counter of note-tuples whose segment id did not fit (sync-free mode)

    @staticmethod
    def _get_segments(aggregate_mode: str, bars=None, beats=None, onsets=None):
        if aggregate_mode in (EmbeddingAggregateModes.BAR_MEAN, EmbeddingAggregateModes.ISOLATED_BAR_MEAN):
            assert bars is not None, f"`bars` should be provided as inputs for aggregate_mode `{aggregate_mode}`"
            return bars
        elif aggregate_mode == EmbeddingAggregateModes.BEAT_MEAN:
            assert beats is not None, f"`beats` should be provided as inputs for aggregate_mode `{aggregate_mode}`"
            return beats
        elif aggregate_mode == EmbeddingAggregateModes.ONSET_MEAN:
            assert onsets is not None, f"`onsets` should be provided as inputs for aggregate_mode `{aggregate_mode}`"
            return onsets
        return None

    def _fused_supported(self) -> bool:
        modes = self.aggregate_mode if isinstance(self.aggregate_mode, list) else [self.aggregate_mode]
        return (isinstance(self.aggregate_mode, list) and self.hierarchical and self.hierarchical_with_context
                and all(m in ("mean",) + SEGMENT_MODES[1:] for m in modes) and self.dim + self.embedding_dim <= 320)

    def forward(self, x: Tensor, mask: Optional[Tensor] = None, x_extra=None, latents=None, bars: Optional[Tensor] = None,
                beats: Optional[Tensor] = None, onsets: Optional[Tensor] = None, deadpan_mask: Optional[Tensor] = None,
                return_embeddings: bool = False, return_attn: bool = False, compute_loss: bool = True,
                z_prior: Optional[List[Tensor]] = None, table_cache: Optional[dict] = None, side_branch: Optional[SideBranch] = None,
                mmd_rows: Optional[List[Optional[Tensor]]] = None,
                **kwargs):
        if latents is not None or not self._fused_supported() or self._mask_bars:
            raise NotImplementedError(
                "scoreperformer_b200.MMDTupleTransformer: the sm_100a path implements hierarchical-with-context levels over "
                "{mean, bar_mean, beat_mean, onset_mean} (every shipped recipe); injected latents / isolated_bar_mean / "
                "non-hierarchical heads are not implemented")
        out_t = super().forward(x=x, mask=mask, x_extra=x_extra, return_embeddings=return_embeddings, return_attn=return_attn,
                                table_cache=table_cache, **kwargs)
        hidden = out_t.hidden_state
        b, t, _ = hidden.shape
        if mask is None:
            mask = torch.ones(b, t, dtype=torch.bool, device=hidden.device)
        mask = mask.contiguous()
        assert not self.deadpan_zero_latent or deadpan_mask is not None

        segs = [self._get_segments(m, bars=bars, beats=beats, onsets=onsets) for m in self.aggregate_mode]
        if self.exact_latent_shapes:
            maxima = torch.stack([s.max() if s is not None else s_zero(hidden) for s in segs]).tolist()   # one host sync
            slots = [2 if s is None else int(m) + 1 for s, m in zip(segs, maxima)]
        else:
            # sync-free mode (CUDA-graph capture): every segment table has `slot_capacity` rows (default T + 4).  Ids count musical
            # time, not notes, so a very sparse window can exceed that; the pooling kernel skips such ids.  They are counted here on
            # the device (`segment_overflow`, read it with TrainStep.segment_overflow_count() or int(...)) so that a run can
            # notice and raise `slot_capacity` (or fall back to exact shapes) instead of diverging silently.
            cap = self.slot_capacity if self.slot_capacity is not None else t + 4
            slots = [2 if s is None else cap for s in segs]
            over = [((s >= cap) & mask).sum() for s in segs if s is not None]
            if over:
                if self.segment_overflow is None or self.segment_overflow.device != hidden.device:
                    self.segment_overflow = torch.zeros((), dtype=torch.int64, device=hidden.device)
                self.segment_overflow += torch.stack(over).sum()
        wb = []
        for m in self.aggregate_mode:
            wb += [self.vae_head[m].linear.weight, self.vae_head[m].linear.bias]
        res = fused.LatentLevelsFn.apply(hidden, mask, tuple(segs), tuple(slots), tuple(self.latent_dim), *wb)
        style = res[0]
        lat_list, lmask_list = list(res[1::2]), list(res[2::2])

        losses: Dict[str, Tensor] = {}
        branch = side_branch if side_branch is not None else SideBranch(hidden.device)
        out_latents = []
        drop_tok = None            # [B, T, n_levels] bool: token-level latent dropout (inclusive over coarser levels)
        level_drops = []
        for i, mode in enumerate(self.aggregate_mode):
            lat, lmask = lat_list[i], lmask_list[i]
            if mode == "mean":       # slot 1 holds the sample; reference shape [B, 1, z] with an all-true mask
                lat, lmask = lat[:, 1:2], torch.ones(b, 1, dtype=torch.bool, device=hidden.device)
            out_latents.append(lat)
            if compute_loss:
                with branch.run(lat, lmask):      # nothing downstream needs these terms before the final loss sum
                    losses[f"MMD/{mode}"] = self.loss_weight * self.criterion(lat, mask=lmask, z=None if z_prior is None else z_prior[i],
                                                                                 rows=None if mmd_rows is None else mmd_rows[i])
                    if self.deadpan_zero_latent:
                        sel = (deadpan_mask[:, None] & lmask)[..., None].to(lat.dtype)       # mmd_transformer.py:268-273
                        denom = sel.sum() * lat.shape[-1]
                        losses[f"MMD/{mode}/deadpan"] = (lat.pow(2) * sel).sum() / denom.clamp(min=1.0)
            # latent dropout (mmd_transformer.py:349-364, 537-542): Bernoulli per valid segment, broadcast to its notes
            p = self.latent_dropout[i]
            if mode != "mean" and self.training and p > 0.:
                seg_drop = (torch.rand(lmask.shape, device=hidden.device) < p) & lmask
                tok_drop = torch.gather(seg_drop, 1, segs[i].clamp(max=lmask.shape[1] - 1))
            else:
                tok_drop = torch.zeros(b, t, dtype=torch.bool, device=hidden.device)
            if self.training and self.inclusive_latent_dropout and level_drops:
                tok_drop = tok_drop | level_drops[-1]
            level_drops.append(tok_drop)

        embeddings = style                                   # already * mask
        if self.training:
            full_embeddings = embeddings
            drop = torch.cat([d[..., None].expand(-1, -1, z) for d, z in zip(level_drops, self.latent_dim)], dim=-1)
            drop = drop & mask[..., None] & ~deadpan_mask[:, None, None] if deadpan_mask is not None else drop & mask[..., None]
            embeddings = embeddings * (~drop)
            drop_mask = drop
        else:
            full_embeddings, drop_mask = embeddings, None

        loss = None
        if compute_loss:
            with branch.run():
                loss = sum(losses.values())
                losses["MMD"] = loss
            if side_branch is None:               # standalone use: rejoin at once; ScorePerformer.forward joins after the decoder
                branch.join(*losses.values())
        return MMDTupleTransformerOutput(hidden_state=hidden, logits=out_t.logits, attentions=None, latents=out_latents,
                                         embeddings=embeddings, full_embeddings=full_embeddings, dropout_mask=drop_mask, loss=loss,
                                         losses=losses)

    # ------------------------------------------------------------------ pure helpers used by the generator (small tensors, torch)
    def embeddings_to_latents(self, embeddings, mask=None, bars=None, beats=None, onsets=None):
        if isinstance(self.aggregate_mode, str):
            seg = self._get_segments(self.aggregate_mode, bars=bars, beats=beats, onsets=onsets)
            return self._embeddings_to_latents(embeddings, self.aggregate_mode, segments=seg, mask=mask)
        parts = embeddings.split(list(self.latent_dim), dim=-1)
        return [self._embeddings_to_latents(parts[i], m, segments=self._get_segments(m, bars=bars, beats=beats, onsets=onsets), mask=mask)
                for i, m in enumerate(self.aggregate_mode)]

    @staticmethod
    def _embeddings_to_latents(embeddings, aggregate_mode, mask=None, segments=None):
        b, t = embeddings.shape[:2]
        if aggregate_mode == EmbeddingAggregateModes.MEAN:
            latents = embeddings.mean(dim=1) if mask is None else embeddings.sum(dim=1) / mask.sum(dim=1)
            return latents.unsqueeze(1)
        if aggregate_mode in SEGMENT_MODES:
            s = int(segments.max()) + 1
            ar = torch.arange(b, device=embeddings.device)[:, None].expand(b, t)
            sums = torch.zeros(b, s, embeddings.shape[-1], dtype=embeddings.dtype, device=embeddings.device)
            sums.index_put_((ar, segments), embeddings, accumulate=True)
            counts = torch.zeros(b, s, dtype=embeddings.dtype, device=embeddings.device)
            counts.index_put_((ar, segments), torch.ones_like(segments, dtype=embeddings.dtype), accumulate=True)
            return sums / counts.clamp(min=1)[..., None]
        return embeddings

    def latents_to_embeddings(self, latents, seq_len, bars=None, beats=None, onsets=None):
        if isinstance(self.aggregate_mode, str):
            seg = self._get_segments(self.aggregate_mode, bars=bars, beats=beats, onsets=onsets)
            return self._latents_to_embeddings(latents, seq_len, self.aggregate_mode, segments=seg)
        embs = [self._latents_to_embeddings(latents[i], seq_len, m, segments=self._get_segments(m, bars=bars, beats=beats, onsets=onsets))
                for i, m in enumerate(self.aggregate_mode)]
        return torch.cat(embs, dim=-1)

    @staticmethod
    def _latents_to_embeddings(latents, seq_len, aggregate_mode, segments=None):
        b, t = latents.shape[0], seq_len
        if aggregate_mode == EmbeddingAggregateModes.MEAN:
            return latents.expand(-1, t, -1)
        if aggregate_mode in SEGMENT_MODES:
            ar = torch.arange(b, device=latents.device)[:, None].expand(b, t)
            return latents[ar, segments]
        return latents


def s_zero(ref: Tensor) -> Tensor:
    return torch.zeros((), dtype=torch.int64, device=ref.device)
