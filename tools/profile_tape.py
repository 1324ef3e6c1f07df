"""A few eval forwards of one dense denoiser (TD_PROFILE_MODEL = diffusion_transformer | latent_diffusion) at batch
TD_PROFILE_BATCH: the target of `ncu -k regex:dense_cluster|dense_tape`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib
import torch
from tinydiff import _lib as L

B = int(os.environ.get("TD_PROFILE_BATCH", "128"))
name = os.environ.get("TD_PROFILE_MODEL", "diffusion_transformer")
dev = L.require_device("cuda:0")
mod = importlib.import_module(f"tinydiff.{name}")
torch.manual_seed(0)
m = mod.NoiseModel().to(dev).eval()
x = torch.randn(B, 20, device=dev)
t = torch.randint(0, 1000, (B,), device=dev)
y = torch.randint(0, 10, (B,), device=dev)
with torch.no_grad():
    for _ in range(6):
        out = m(x, t, y)
torch.cuda.synchronize()
print("ok", float(out.abs().max()))
