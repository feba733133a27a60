"""Worker for test_trainer_nccl_world2 (torch.distributed.run, NCCL, one process per GPU): DataParallelTrainer with a
DistributedSampler -- the two ranks see disjoint shards, hold identical parameters after an epoch, rank 0 alone writes the
checkpoint, and both ranks resume from it to the same state."""
import os
import sys
import tempfile

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests import parity  # noqa: E402
from scoreperformer_b200.synthetic import SyntheticDataset, collate_rows  # noqa: E402
from scoreperformer_b200.trainer import DataParallelTrainer, steps_per_epoch  # noqa: E402


def build(seed=0):
    torch.manual_seed(seed)
    model = parity.build_model(dropout=False, device="cuda").train()
    model.perf_decoder.label_fields = (3, 5, 10, 11)
    return model


def main():
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    rank, world = dist.get_rank(), dist.get_world_size()
    data = SyntheticDataset(32, 48, seed=7)
    out_dir = os.environ["SPB_TEST_OUT"]

    tr = DataParallelTrainer(build(), data, collate_rows, batch_size=4, lr=1e-3, lr_gamma=0.9, output_dir=out_dir, seed=3, use_graph=True)
    loader = tr.build_dataloader(data)
    tr.sampler.set_epoch(0)
    mine = torch.tensor(list(iter(tr.sampler)), device="cuda")
    both = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(both, mine)
    assert len(set(torch.cat(both).tolist())) == world * len(mine) == 32, "ranks must see disjoint shards covering the epoch"
    assert len(loader) == steps_per_epoch(32, 4, world) == 4

    torch.manual_seed(11)                      # MMD prior samples
    logs = tr.fit(epochs=1)
    assert [e["step"] for e in logs] == [1, 2, 3, 4] and all(e["loss"] == e["loss"] for e in logs)
    flat = tr.step.flat_param.clone()
    others = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(others, flat)
    assert all(torch.equal(o, others[0]) for o in others), "parameters diverged between ranks"
    path = tr.save_checkpoint()
    assert (path is not None) == (rank == 0)
    path = os.path.join(out_dir, "checkpoint_4.pt")
    assert os.path.exists(path)
    torch.manual_seed(12)
    more = tr.fit(epochs=2)                     # epoch 1 on top
    ref_after = tr.step.flat_param.clone()
    ref_loss = [e["loss"] for e in more if e["step"] > 4]
    tr.close()

    tr2 = DataParallelTrainer(build(seed=99), data, collate_rows, batch_size=4, lr=1e-3, lr_gamma=0.9, output_dir=out_dir, seed=3, use_graph=True)
    tr2.load_checkpoint(path)
    assert tr2.epoch == 1 and tr2.global_step == 4
    torch.manual_seed(12)
    again = tr2.fit(epochs=2)
    got_loss = [e["loss"] for e in again]
    assert len(got_loss) == len(ref_loss) == 4
    for a, b in zip(got_loss, ref_loss):
        assert abs(a - b) <= 2e-3 * abs(b), (got_loss, ref_loss)
    dev = float((tr2.step.flat_param - ref_after).abs().max() / ref_after.abs().max())
    assert dev < 2e-3, dev
    tr2.close()
    dist.barrier()
    print("TRAINER-NCCL-OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
