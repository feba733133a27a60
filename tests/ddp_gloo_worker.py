"""Worker for test_gradient_buckets_gloo_world2 (launched with torch.distributed.run, gloo backend, CPU)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from scoreperformer_b200.parallel import GradientBuckets, default_buckets  # noqa: E402
from tests import parity  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    model = parity.build_model()
    names = default_buckets(model)
    flat = [n for b in names for n in b]
    uniq = {id(p) for p in model.parameters()}
    assert len(flat) == len(uniq) == 236, (len(flat), len(uniq))           # every unique tensor in exactly one bucket
    assert all(".token_emb.embs." in n for n in names[-1])                  # tied tables reduce last
    buckets = GradientBuckets(model, overlap=False)
    # fake backward: every parameter gets gradient (rank + 1) * f(param) through autograd so the hooks fire
    loss = sum(((rank + 1.0) * p * torch.full_like(p, 0.5)).sum() for p in model.parameters())
    loss.backward()
    buckets.sync_gradients()
    expect = 0.5 * sum(r + 1.0 for r in range(world)) / world
    for n, p in model.named_parameters():
        assert torch.allclose(p.grad, torch.full_like(p, expect)), n
    # second step re-arms the buckets
    model.zero_grad(set_to_none=True)
    loss = sum(((rank + 2.0) * p).sum() for p in model.parameters())
    loss.backward()
    buckets.sync_gradients()
    expect = sum(r + 2.0 for r in range(world)) / world
    for n, p in model.named_parameters():
        assert torch.allclose(p.grad, torch.full_like(p, expect)), n
    print("DDP-OK", rank, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
