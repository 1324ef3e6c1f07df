"""Short profiling target: a few eager eval forwards of the benchmark model at the benchmark batch
(the same kernels, buffers and order as one reverse step of bench.py), followed by p_sample.
Used under ncu (launch list / --set full); never a source of benchmark numbers."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinydiff import _lib as L
from tinydiff.conditional_diffusion import ForwardProcess, NoiseModel
from tinydiff.process import ReverseLoop

B = int(os.environ.get("TD_PROFILE_BATCH", "128"))
iters = int(os.environ.get("TD_PROFILE_ITERS", "3"))
dev = L.require_device("cuda:0")
torch.manual_seed(0)
model = NoiseModel().to(dev).eval()
fp = ForwardProcess()
eng = model.engine(B, dev)
eng.refresh_weights()
eng.x_in.copy_(torch.randn(B, 1, 28, 28))
eng.y_in.copy_(torch.randint(0, 10, (B,)))
eng.use_t_dev = True
eng.prepare_sampler_embed()
loop = ReverseLoop(fp, eng.x_in, eng.eps, eng.t_dev, eng.launch, use_graph=False)
torch.cuda.synchronize()
loop.run(seed=3, steps=iters)
torch.cuda.synchronize()
print("launches per reverse step:", eng.num_launches() + 2, "finite:", bool(torch.isfinite(eng.x_in).all()))
