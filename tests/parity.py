"""Shared parity harness: build the sm_100a model and the CPU oracle on identical weights / batches and compare.

TEST INFRASTRUCTURE: this module (not the product package) is the only place that wires the oracle to the CUDA path.
"""
from __future__ import annotations

import copy
import os
import sys
from typing import Dict, List

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import model_oracle as mo  # noqa: E402
from weights import fill_model_  # noqa: E402

from scoreperformer_b200.models import ScorePerformer  # noqa: E402
from scoreperformer_b200.recipes import default_model_config  # noqa: E402
from scoreperformer_b200.synthetic import make_batch  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def build_model(dropout: bool = False, seed_weights: int = 0, device: str = "cpu") -> ScorePerformer:
    cfg = default_model_config(dropout=dropout)
    model = ScorePerformer.init(copy.deepcopy(cfg))
    fill_model_(model, seed_weights)
    return model.to(device)


def oracle_state(model: torch.nn.Module, requires_grad: bool = True) -> Dict[str, torch.Tensor]:
    """CPU fp32 copy of the state_dict that preserves parameter tying (aliases map to ONE tensor object)."""
    sd, first = {}, {}
    for k, v in model.state_dict(keep_vars=True).items():
        ptr = v.data_ptr()
        if ptr in first:
            sd[k] = sd[first[ptr]]
            continue
        first[ptr] = k
        t = v.detach().cpu().clone()
        if requires_grad and isinstance(v, torch.nn.Parameter):
            t.requires_grad_(True)
        sd[k] = t
    return sd


def oracle_spec(model: ScorePerformer) -> mo.OracleSpec:
    return mo.spec_from_config(default_model_config(dropout=False))


def golden(name: str):
    return np.load(os.path.join(GOLDEN_DIR, name), allow_pickle=False)


def z_from_golden(g) -> List[torch.Tensor]:
    return [torch.from_numpy(g[f"z{i}"]) for i in range(4)]


def cosine_distance(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().flatten(), b.double().flatten()
    return float(1 - torch.dot(a, b) / (a.norm() * b.norm()).clamp(min=1e-30))


def run_product_step(model: ScorePerformer, batch, z):
    """One training forward + backward of the CUDA model; returns the outputs object (grads live on the parameters)."""
    dev = next(model.parameters()).device
    model.train()
    model.zero_grad(set_to_none=True)
    model.z_prior = [t.to(dev) for t in z]
    out = model(**{k: v.to(dev) for k, v in batch.items()})
    out.loss.backward()
    return out


def run_oracle_step(model: ScorePerformer, batch, z):
    sd = oracle_state(model)
    spec = oracle_spec(model)
    out = mo.scoreperformer_forward(sd, batch, spec, [t.clone() for t in z])
    out["loss"].backward()
    return out, sd


def compare_step(model, batch, z, loss_rtol=2e-2, cos_tol=5e-3, verbose=True):
    """north_star tolerances: losses/logits within 2e-2 relative (bf16), gradient cosine distance <= 5e-3."""
    out = run_product_step(model, batch, z)
    ref, sd = run_oracle_step(model, batch, z)
    report = {}
    for key, val in ref["losses"].items():
        got = float(out.losses[key])
        want = float(val)
        report[f"loss/{key}"] = (got, want)
        assert abs(got - want) <= loss_rtol * max(abs(want), 1e-2), f"loss {key}: {got} vs oracle {want}"
    got, want = float(out.loss), float(ref["loss"])
    assert abs(got - want) <= loss_rtol * abs(want), f"total loss {got} vs oracle {want}"
    # hidden states / embeddings
    for name, got_t in (("score_hidden", out.score_encoder.hidden_state), ("perf_hidden", out.perf_encoder.hidden_state),
                        ("embeddings", out.perf_encoder.embeddings), ("dec_hidden", out.perf_decoder.hidden_state)):
        want_t = ref[name]
        err = float((got_t.detach().float().cpu() - want_t.detach()).abs().max() / want_t.detach().abs().max())
        report[f"relerr/{name}"] = err
        assert err < 6e-2, f"{name}: max rel err {err}"
    # pooling membership is bit-exact: the latents' validity pattern must match exactly
    for lat, lat_ref in zip(out.perf_encoder.latents, ref["latents"]):
        assert lat.shape == lat_ref.shape, (lat.shape, lat_ref.shape)
        assert torch.equal((lat.detach().cpu() != 0).any(-1), (lat_ref.detach() != 0).any(-1)), "latent validity differs"
    # gradients
    worst = (0.0, None)
    seen = set()
    params = dict(model.state_dict(keep_vars=True))
    for k, p in params.items():
        if not isinstance(p, torch.nn.Parameter) or p.data_ptr() in seen:
            continue
        seen.add(p.data_ptr())
        g_ref = sd[k].grad
        assert p.grad is not None, f"no gradient for {k}"
        assert g_ref is not None, f"oracle has no gradient for {k}"
        g_got = p.grad.detach().float().cpu()
        if float(g_ref.norm()) == 0.0:      # e.g. index rows of an unlabelled field: exactly no gradient in the reference
            assert float(g_got.abs().max()) == 0.0, f"{k}: reference gradient is exactly zero, got {float(g_got.abs().max())}"
            continue
        cd = cosine_distance(g_got, g_ref)
        norm_ratio = float(g_got.norm() / g_ref.norm())
        assert 0.9 < norm_ratio < 1.1, f"grad norm ratio {norm_ratio:.3f} for {k}"
        if cd > worst[0]:
            worst = (cd, k)
        assert cd <= cos_tol, f"grad cosine distance {cd:.2e} > {cos_tol} for {k}"
    report["worst_grad_cosine_distance"] = worst
    if verbose:
        for k, v in report.items():
            print(f"  {k}: {v}")
    return report


def run_smoke():
    """__graft_entry__.smoke(): one tiny forward+backward on cuda:0 checked against the oracle."""
    assert torch.cuda.is_available(), "smoke() needs a CUDA device"
    torch.manual_seed(0)
    model = build_model(dropout=False, device="cuda:0")
    batch = make_batch(2, 48, seed=1234)
    z = [torch.randn(256, d) for d in (32, 20, 8, 4)]
    rep = compare_step(model, batch, z, verbose=False)
    print("smoke OK: total loss (cuda, oracle) =", rep["loss/MMD"], "worst grad cos dist =", rep["worst_grad_cosine_distance"])
