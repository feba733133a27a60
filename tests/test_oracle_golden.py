"""CPU: the oracle (oracle/model_oracle.py) is pinned to vectors produced by the UNMODIFIED reference (oracle/gen_golden.py)."""
import numpy as np
import pytest
import torch

from tests import parity
import model_oracle as mo


def test_known_answer_vectors():
    """SURVEY Appendix D vectors recomputed from the reference's pure functions."""
    g = parity.golden("known_answers.npz")
    emb, seg = torch.from_numpy(g["seg_emb"]), torch.from_numpy(g["seg_ids"])
    pooled, counts = mo.segment_mean(emb, seg)
    assert np.allclose(pooled.numpy(), g["seg_latents"])
    back = pooled[torch.arange(1)[:, None].expand(1, 6), seg]
    assert np.allclose(back.numpy(), g["seg_back"])
    x, y = torch.from_numpy(g["mmd_x"]), torch.from_numpy(g["mmd_y"])
    assert abs(float(mo.mmd(x, y)) - float(g["mmd_value"])) < 1e-6
    k = torch.exp(-((x[:, None] - y[None]) ** 2).mean(-1) / 2)
    assert np.allclose(k.numpy(), g["mmd_kernel"], atol=1e-6)
    slopes = torch.tensor([0.25, 0.0625, 0.015625, 0.00390625])
    assert np.allclose(mo.alibi_bias(2, 4, 2, slopes).numpy(), g["alibi"], atol=1e-7)
    assert np.allclose(mo.class_weights([0.9, 0.09, 0.01]), g["class_weights"], atol=1e-7)


@pytest.mark.parametrize("name", ["train_b2_t48.npz", "train_b3_t33.npz"])
def test_oracle_training_step_matches_reference(name):
    """Losses, hidden states, logits and gradients of one training step, fp32 round-off."""
    g = parity.golden(name)
    model = parity.build_model(dropout=False, device="cpu")      # product module tree on CPU: parameters only, no forward
    batch = parity.make_batch(int(g["B"]), int(g["T"]), seed=int(g["seed"]))
    out, sd = parity.run_oracle_step(model, batch, parity.z_from_golden(g))
    assert abs(float(out["loss"]) - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    assert list(out["losses"].keys()) == [str(k) for k in g["loss_keys"]]
    for key, want in zip(g["loss_keys"], g["loss_vals"]):
        assert abs(float(out["losses"][str(key)]) - want) < 1e-4 * max(1.0, abs(want)), key
    for name_ in ("score_hidden", "perf_hidden", "embeddings", "dec_hidden"):
        assert np.abs(out[name_].detach().numpy() - g[name_]).max() < 1e-4, name_
    for key in ("Velocity", "Tempo", "RelOnsetDev", "RelPerfDuration", "Bar", "NotesInOnset"):
        assert np.abs(out["logits"][key].detach().numpy() - g[f"logits/{key}"]).max() < 2e-4, key
    for i in range(4):
        assert out["latents"][i].shape == g[f"latents{i}"].shape            # S_l = max id + 1: membership bit-exact
        assert np.abs(out["latents"][i].detach().numpy() - g[f"latents{i}"]).max() < 1e-4
    for k in g.files:
        if k.startswith("grad/"):
            ref = g[k]
            got = sd[k[5:]].grad.numpy()
            assert np.abs(got - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max()), k
    norms = dict(zip([str(k) for k in g["grad_norm_keys"]], g["grad_norm_vals"]))
    for k, want in norms.items():
        got = float(sd[k].grad.norm()) if sd[k].grad is not None else 0.0
        assert abs(got - want) <= 1e-3 * max(want, 1e-3), k


def test_oracle_cached_rendering_matches_reference():
    """Greedy note-by-note rendering (cached and uncached) reproduces the reference's `unmask_tokens` tokens exactly."""
    g = parity.golden("render_t24.npz")
    model = parity.build_model(dropout=False, device="cpu")
    sd = parity.oracle_state(model, requires_grad=False)
    spec = parity.oracle_spec(model)
    tokens_in = torch.from_numpy(g["tokens_in"])
    batch = parity.make_batch(1, int(g["T"]), seed=int(g["seed"]), full_length=True, deadpan_last=False)
    score = torch.from_numpy(g["score_embeddings"])
    style = torch.from_numpy(g["perf_embeddings"])
    for use_cache in (True, False):
        out = mo.render_greedy(sd, spec, tokens_in, batch["masked_perf"], score, style, use_cache=use_cache)
        assert torch.equal(out, torch.from_numpy(g["tokens_out"])), f"use_cache={use_cache}"
    with torch.no_grad():
        sh = mo.encoder_forward(sd, "score_encoder", batch["score"], batch["score_mask"], list(spec.num_score_tokens), spec.depth_score, spec)
        enc = mo.perf_encoder_forward(sd, batch, spec, None, training=False, compute_loss=False)
    assert np.abs(sh.numpy() - g["score_embeddings"]).max() < 1e-4
    assert np.abs(enc["embeddings"].numpy() - g["perf_embeddings"]).max() < 1e-4
