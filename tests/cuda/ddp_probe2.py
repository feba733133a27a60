import os, sys, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import parity
from scoreperformer_b200.train_step import TrainStep

dist.init_process_group("nccl")
rank = dist.get_rank()
torch.cuda.set_device(rank)
same = {k: v.cuda() for k, v in parity.make_batch(2, 64, seed=5).items()}


def run(group, overlap):
    os.environ["SPB_DDP_OVERLAP"] = "1" if overlap else "0"
    torch.manual_seed(0)
    model = parity.build_model(dropout=False, device="cuda").train()
    model.perf_decoder.label_fields = (3, 5, 10, 11)
    ts = TrainStep(model, lr=0.0, use_graph=False, process_group=group)
    if group is None:
        ts.world, ts.overlap = 1, False
    torch.manual_seed(1)
    local = []
    orig = ts._update

    def upd():
        torch.cuda.synchronize()
        local.append(ts.flat_grad.clone())
        orig()
    ts._update = upd
    for i in range(2):
        loss = ts.step(same)
    torch.cuda.synchronize()
    return local[-1], ts.flat_grad.clone() / ts.world, float(loss), ts


g_single_local, g_single, l0, _ = run(None, False)
for overlap in (False, True):
    loc, avg, l1, ts = run(dist.group.WORLD, overlap)
    d1 = float((loc - g_single).abs().max() / g_single.abs().max())
    d2 = float((avg - g_single).abs().max() / g_single.abs().max())
    print(f"rank {rank} overlap {overlap}: loss {l0:.6f} vs {l1:.6f}; local-before-reduce vs single {d1:.2e}; averaged vs single {d2:.2e}", flush=True)
    ts.close()
dist.barrier()
dist.destroy_process_group()
