"""One small eager training step for compute-sanitizer:
   PYTORCH_NO_CUDA_MEMORY_CACHING=1 compute-sanitizer --tool initcheck --print-limit 40 python tests/cuda/initcheck_step.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import parity
from scoreperformer_b200.train_step import TrainStep

torch.manual_seed(0)
model = parity.build_model(dropout=False, device="cuda").train()
model.perf_decoder.label_fields = (3, 5, 10, 11)
ts = TrainStep(model, lr=1e-3, use_graph=False)
batch = {k: v.cuda() for k, v in parity.make_batch(2, 64, seed=5).items()}
for i in range(2):
    print("step", i, float(ts.step(batch)), flush=True)
torch.cuda.synchronize()
