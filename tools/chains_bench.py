"""1000-step sampler at batch 128 with TD_SAMPLE_CHAINS sub-batches (process.SamplerChains): device-resident timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinydiff import _lib as L
from tinydiff.conditional_diffusion import ForwardProcess, NoiseModel
from tinydiff.diffusion import sampler_plan
dev = L.require_device("cuda:0")
B = int(os.environ.get("TD_PROFILE_BATCH", "128"))
torch.manual_seed(0)
model = NoiseModel().to(dev).eval()
fp = ForwardProcess(1000)
g = torch.Generator().manual_seed(1)
xT = torch.randn(B, 1, 28, 28, generator=g).to(dev)
y = torch.randint(0, 10, (B,), generator=g).to(dev)
plan = sampler_plan(model, fp, dev, B)
def step():
    plan.load(xT, y)
    plan.run(z=None, seed=7)
for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"chains={len(plan.engs)} B={B}: {ms:.1f} ms per 1000 steps = {ms:.1f} us/step -> {B / ms * 1e3:.1f} samples/s  "
      f"(launches/step {plan.loop.launches_per_step}) finite={bool(torch.isfinite(plan.result()).all())}")
