"""The default recipe's `model:` section as data (recipes/scoreperformer/base.yaml:68-192 resolved over default.yaml),
plus the SPMupleWindow vocabulary injected the way `ScorePerformer.inject_data_config` does
(models/scoreperformer/model.py:374-394).  Used by bench.py / tests where the reference's recipes directory is absent;
`config.load_recipe` loads the real YAML files unchanged when they are available (tests/test_config.py checks both agree).
"""
from __future__ import annotations

import copy
from typing import Dict, Optional

from .config import DictConfig, wrap
from .synthetic import DIRECTION_CLASSES, PERF_SIZES, SCORE_KEYS

_ATTENTION = {"dim_head": 64, "one_kv_head": True, "dropout": 0.1, "alibi_pos_bias": True, "alibi_learned": True}
_FEED_FORWARD = {"mult": 4, "glu": True, "swish": True, "dropout": 0.1}


def _token_embeddings(target: str) -> Dict:
    cfg = {"_target_": target, "emb_dims": 128, "mode": "cat", "emb_norm": True, "discrete": False, "continuous": True,
           "continuous_dense": True, "discrete_ids": [0, 1, 2, 3], "tie_keys": None}
    if target == "multi-seq":
        cfg["multiseq_mode"] = "post-cat"
    return cfg


def _transformer(target: str, depth: int) -> Dict:
    return {"_target_": target, "depth": depth, "heads": 4, "attention": copy.deepcopy(_ATTENTION),
            "feed_forward": copy.deepcopy(_FEED_FORWARD)}


def default_model_config(num_tokens: Optional[Dict[str, int]] = None, direction_classes: Optional[Dict[str, int]] = None,
                         dropout: bool = True) -> DictConfig:
    """Resolved `model:` node of recipes/scoreperformer/base.yaml with data-dependent fields injected.

    `dropout=False` zeroes the four dropouts (attention, feed-forward, latent, classifier) for parity runs."""
    import numpy as np
    num_tokens = dict(num_tokens or PERF_SIZES)
    direction_classes = dict(direction_classes or DIRECTION_CLASSES)
    score_tokens = {k: v for k, v in num_tokens.items() if k in SCORE_KEYS}
    token_values = {k: [0.0, 0.0, 0.0, 0.0] + np.linspace(0.0, 1.0, v - 4).tolist() for k, v in num_tokens.items()}

    def with_values(te: Dict, keys) -> Dict:
        te = dict(te)
        te["token_values"] = {k: list(token_values[k]) for k in keys}
        return te

    cfg = {
        "_name_": "ScorePerformer", "_version_": "v0.4.4", "dim": 256, "tie_token_emb": True, "mode": "mixlm",
        "num_tokens": num_tokens, "num_score_tokens": score_tokens,
        "score_encoder": {
            "token_embeddings": with_values(_token_embeddings("simple"), score_tokens),
            "emb_norm": True, "emb_dropout": 0, "use_abs_pos_emb": False, "transformer": _transformer("encoder", 2),
        },
        "perf_encoder": {
            "token_embeddings": with_values(_token_embeddings("simple"), num_tokens),
            "emb_norm": True, "emb_dropout": 0, "use_abs_pos_emb": False,
            "latent_dim": [32, 20, 8, 4], "aggregate_mode": ["mean", "bar_mean", "beat_mean", "onset_mean"],
            "latent_dropout": [0.0, 0.1, 0.2, 0.4], "hierarchical": True, "inclusive_latent_dropout": True,
            "deadpan_zero_latent": True, "loss_weight": 1.0, "transformer": _transformer("encoder", 4),
        },
        "perf_decoder": {
            "token_embeddings": with_values(_token_embeddings("multi-seq"), num_tokens),
            "emb_norm": True, "emb_dropout": 0, "use_abs_pos_emb": False, "context_emb_mode": "cat",
            "style_emb_dim": [32, 20, 8, 4], "style_emb_mode": "adanorm", "transformer": _transformer("decoder", 4),
            "lm_head": {"_target_": "lm-tied"}, "regression_head": None,
        },
        "classifiers": {
            "classifier": {"hidden_dims": [], "dropout": 0.2}, "loss_weight": 1.0, "weighted_classes": True, "detach_inputs": True,
            "num_classes": direction_classes,
            "class_samples": {k: [1.0 / v] * v for k, v in direction_classes.items()},
        },
    }
    if not dropout:
        for stack in ("score_encoder", "perf_encoder", "perf_decoder"):
            cfg[stack]["transformer"]["attention"]["dropout"] = 0.0
            cfg[stack]["transformer"]["feed_forward"]["dropout"] = 0.0
        cfg["perf_encoder"]["latent_dropout"] = [0.0, 0.0, 0.0, 0.0]
        cfg["classifiers"]["classifier"]["dropout"] = 0.0
    return wrap(cfg)
