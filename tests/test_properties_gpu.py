"""Size-independent properties of the sm_100a path at BASELINE.json's full sizes (configs[1]: 64 x 512, configs[3]: 16 x 2048),
where the CPU oracle would take too long: pooling membership, causality, padding invariance, linearity, determinism."""
import pytest
import torch

from tests import parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model():
    m = parity.build_model(dropout=False, device="cuda")
    m.perf_decoder.label_fields = (3, 5, 10, 11)
    return m


def _batch(B, T, seed):
    return {k: v.cuda() for k, v in parity.make_batch(B, T, seed=seed).items()}


@pytest.mark.parametrize("B,T", [(64, 512), (16, 2048)])
def test_pooling_membership_bit_exact_at_full_size(model, B, T):
    """Segment counts of every level equal an integer bincount of the segment ids (bit-exact), latents of empty slots are zero."""
    from scoreperformer_b200 import kernels as K
    b = _batch(B, T, seed=11)
    hidden = torch.randn(B, T, 256, device="cuda")
    for seg_name, z in (("bars", 20), ("beats", 8), ("onsets", 4)):
        seg = b[seg_name]
        S = int(seg.max()) + 1
        style = torch.zeros(B, T, 64, device="cuda")
        W, bias = torch.randn(z, 256, device="cuda") * 0.1, torch.zeros(z, device="cuda")
        lat, lmask, pooled, counts = K.latent_level_fwd(hidden, style, b["perf_mask"], seg, W, bias, 0, S, z)
        want = torch.zeros(B, S, dtype=torch.int64, device="cuda").scatter_add_(1, seg, torch.ones_like(seg))
        assert torch.equal(counts.view(B, S).long(), want), seg_name
        assert not lmask[:, 0].any(), "the padding segment (id 0) is never valid"
        assert float(lat[~lmask].abs().max()) == 0.0
        # every valid note reads back exactly the latent of its own segment
        emb = style[..., :z]
        gathered = torch.gather(lat, 1, seg[..., None].expand(-1, -1, z)) * b["perf_mask"][..., None]
        assert torch.equal(emb, gathered)


def test_full_size_step_is_finite_and_decoder_is_causal(model):
    """configs[1]: one training forward+backward is finite; changing the notes after position p leaves decoder states <= p-1
    untouched (causal attention + MixedLM shift), bit for bit."""
    B, T, p = 8, 512, 300
    b = _batch(B, T, seed=21)
    model.eval()
    with torch.no_grad():
        enc = model.forward_encoders(perf=b["perf"], perf_mask=b["perf_mask"], score=b["score"], score_mask=b["score_mask"],
                                     bars=b["bars"], beats=b["beats"], onsets=b["onsets"], deadpan_mask=b["deadpan_mask"], compute_loss=False)
        kw = dict(mask=b["perf_mask"], style_embeddings=enc.perf_embeddings, context=enc.score_embeddings, return_embeddings=True)
        h1 = model.perf_decoder(b["perf"], seq_masked=b["masked_perf"], **kw).hidden_state
        perf2, masked2 = b["perf"].clone(), b["masked_perf"].clone()
        perf2[:, p:, 3] = torch.randint(4, 132, perf2[:, p:, 3].shape, device="cuda") * b["perf_mask"][:, p:]
        masked2[:, p + 1:, 1] = 7
        h2 = model.perf_decoder(perf2, seq_masked=masked2, **kw).hidden_state
    assert torch.equal(h1[:, :p - 1], h2[:, :p - 1]), "decoder position i must not see notes > i+1"
    assert not torch.equal(h1[:, p:], h2[:, p:])
    model.train()
    model.zero_grad(set_to_none=True)
    out = model(**_batch(64, 512, seed=22))
    out.loss.backward()
    assert torch.isfinite(out.loss)
    total = sum(float(q.grad.double().pow(2).sum()) for q in model.parameters())
    assert total > 0 and total == total


def test_padding_content_is_invisible(model):
    """Tokens / segments stored at padded positions never influence valid outputs or the loss (encoders are non-causal)."""
    b = _batch(16, 512, seed=31)
    model.eval()
    pad = ~b["perf_mask"]
    b2 = {k: v.clone() for k, v in b.items()}
    b2["perf"][pad] = 5
    b2["score"][pad] = 5
    b2["masked_perf"][pad] = 5
    with torch.no_grad():
        o1 = model.forward_encoders(perf=b["perf"], perf_mask=b["perf_mask"], score=b["score"], score_mask=b["score_mask"],
                                    bars=b["bars"], beats=b["beats"], onsets=b["onsets"], deadpan_mask=b["deadpan_mask"], compute_loss=False)
        o2 = model.forward_encoders(perf=b2["perf"], perf_mask=b["perf_mask"], score=b2["score"], score_mask=b["score_mask"],
                                    bars=b["bars"], beats=b["beats"], onsets=b["onsets"], deadpan_mask=b["deadpan_mask"], compute_loss=False)
    m = b["perf_mask"]
    assert torch.equal(o1.score_embeddings[m], o2.score_embeddings[m])
    # segment pooling sums with fp32 atomics, so the latents are reproducible to rounding, not bit for bit
    assert float((o1.perf_embeddings - o2.perf_embeddings).abs().max()) < 1e-5


def test_gemm_linearity_and_determinism_full_size():
    """C = A B^T is linear in A and bit-reproducible run to run at the C2 shapes (no atomics on the non-split path)."""
    from scoreperformer_b200 import kernels as K
    torch.manual_seed(0)
    a = torch.randn(32768, 256, device="cuda").bfloat16()
    w = torch.randn(2048, 256, device="cuda").bfloat16()
    c1 = K.gemm(a, w, out_dtype=torch.float32)
    c2 = K.gemm(a, w, out_dtype=torch.float32)
    assert torch.equal(c1, c2)
    c4 = K.gemm((a.float() * 4).bfloat16(), w, out_dtype=torch.float32)      # exact power-of-two scaling
    assert torch.equal(c4, c1 * 4)
    # split-K wgrad: sum over row blocks equals the whole (fp32 atomics: tolerance, not bit-exact)
    dy = torch.randn(32768, 384, device="cuda").bfloat16()
    full = K.gemm(dy, a, trans_a=True, trans_b=True, out_dtype=torch.float32, split_k=0)
    parts = K.gemm(dy[:16384], a[:16384], trans_a=True, trans_b=True, out_dtype=torch.float32, split_k=0) + \
        K.gemm(dy[16384:], a[16384:], trans_a=True, trans_b=True, out_dtype=torch.float32, split_k=0)
    assert float((full - parts).abs().max()) <= 1e-3 * float(full.abs().max())


def test_attention_rows_are_convex_combinations_full_size():
    """With V = const per batch the attention output equals that constant for every valid query (softmax rows sum to 1),
    at T = 2048 with ragged key padding, for the mma.sync and the tcgen05 forward."""
    from scoreperformer_b200 import kernels as K
    B, T, H = 4, 2048, 4
    qkv = torch.randn(B * T, 384, device="cuda").bfloat16()
    const = torch.tensor([0.5, -1.0, 2.0, 0.25], device="cuda").bfloat16()
    qkv.view(B, T, 384)[:, :, 320:] = const[:, None, None]
    lengths = torch.tensor([2048, 1500, 1024, 2001], device="cuda")
    mask = torch.arange(T, device="cuda")[None] < lengths[:, None]
    ls = torch.log(torch.tensor([0.25, 0.0625, 0.015625, 0.0039], device="cuda"))
    for impl in ("mma", "tcgen05"):
        for causal in (False, True):
            out, _, _ = K.attention_fwd(qkv, mask, ls, B, T, H, causal, 0.0, 0, impl=impl)
            want = const.float()[:, None, None].expand(B, T, 256)
            assert float((out.view(B, T, 256).float() - want).abs().max()) < 2e-2, (impl, causal)
