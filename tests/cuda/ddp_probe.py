"""One TrainStep configuration under NCCL with progress prints (debugging aid): GRAPH=0/1 OVERLAP=0/1 INGRAPH=0/1
   timeout 90 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/cuda/ddp_probe.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import parity  # noqa: E402
from scoreperformer_b200.train_step import TrainStep  # noqa: E402

dist.init_process_group("nccl")
rank = dist.get_rank()
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
os.environ["SPB_DDP_OVERLAP"] = os.environ.get("OVERLAP", "1")
os.environ["SPB_DDP_GRAPH"] = os.environ.get("INGRAPH", "1")
torch.manual_seed(0)
model = parity.build_model(dropout=False, device="cuda").train()
model.perf_decoder.label_fields = (3, 5, 10, 11)
ts = TrainStep(model, lr=1e-3, use_graph=os.environ.get("GRAPH", "1") == "1", process_group=dist.group.WORLD)
batch = {k: v.cuda() for k, v in parity.make_batch(2, 64, seed=10 + rank).items()}
for i in range(6):
    loss = ts.step(batch)
    torch.cuda.synchronize()
    print(f"rank {rank} step {i} loss {float(loss):.4f} buckets {len(ts._buckets)} launched {ts._next_bucket}", flush=True)
p = ts.flat_param.clone()
ref = p.clone()
dist.broadcast(ref, src=0)
print(f"rank {rank} params equal rank 0: {torch.equal(p, ref)}", flush=True)
ts.close()
dist.barrier()
print(f"rank {rank} closing", flush=True)
dist.destroy_process_group()
print(f"rank {rank} done", flush=True)
