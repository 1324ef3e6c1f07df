"""Short profiling target: a few eager train steps of the benchmark model at the benchmark batch
(same kernels / buffers / order as bench.py's train section).  Used under ncu; never a source of numbers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinydiff import _lib as L
from tinydiff.conditional_diffusion import ForwardProcess, NoiseModel
from tinydiff.train import TrainStep
B = int(os.environ.get("TD_PROFILE_BATCH", "128"))
iters = int(os.environ.get("TD_PROFILE_ITERS", "2"))
dev = L.require_device("cuda:0")
torch.manual_seed(0)
model = NoiseModel().to(dev).train()
ts = TrainStep(model, ForwardProcess(), B, dev, use_graph=False)
x0 = torch.rand(B, 1, 28, 28) * 2 - 1
y = torch.randint(0, 10, (B,))
for _ in range(iters):
    loss = ts(x0, y)
torch.cuda.synchronize()
print("loss", float(loss), "launch plan", ts.eng.num_launches())
