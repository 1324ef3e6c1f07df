"""Profiling target: two eager (no CUDA graph) fused train steps of the LAION latent UNet (config 5) at batch TD_PROFILE_BATCH
(default 256).  Used under `ncu --metrics gpu__time_duration.sum` for the per-kernel launch list; never a source of bench numbers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinydiff import _lib as L
from tinydiff.conditional_diffusion_laion import ForwardProcess, NoiseModel
from tinydiff.train import TrainStep
B = int(os.environ.get("TD_PROFILE_BATCH", "256"))
dev = L.require_device("cuda:0")
torch.manual_seed(0)
m = NoiseModel().to(dev).train()
ts = TrainStep(m, ForwardProcess(), B, dev, use_graph=False)
x0 = torch.randn(B, 4, 32, 32)
text = torch.randn(B, 768)
for _ in range(int(os.environ.get("TD_PROFILE_STEPS", "2"))):
    loss = ts(x0, text)
torch.cuda.synchronize()
print("loss", float(loss))
