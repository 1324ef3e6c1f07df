"""Short profiling target for the r02 kernels outside the conv UNet: the fused dense tape (latent MLP / DiT eval forward at batch
128) and the vae_laion kernels (encode + decode of one 256x256 image).  Used under ncu; never a source of benchmark numbers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinydiff import _lib as L
dev = L.require_device("cuda:0")
torch.manual_seed(0)
import tinydiff.latent_diffusion as LD
import tinydiff.diffusion_transformer as DT
from tinydiff.vae_laion import VAE
for mod in (LD, DT):
    m = mod.NoiseModel().to(dev).eval()
    x = torch.randn(128, 20, device=dev)
    t = torch.randint(0, 1000, (128,), device=dev)
    y = torch.randint(0, 10, (128,), device=dev)
    with torch.no_grad():
        for _ in range(3):
            out = m(x, t, y)
    torch.cuda.synchronize()
    print(mod.__name__, float(out.abs().mean()))
from oracle.fixtures import init_state_dict
vae = VAE()
vae.load_state_dict(init_state_dict("vae_laion"))
vae = vae.to(dev).eval()
img = torch.rand(1, 3, 256, 256, device=dev)
for _ in range(2):
    mu, lv = vae.encode(img)
    rec = vae.decode(mu)
torch.cuda.synchronize()
print("vae_laion", float(mu.abs().mean()), float(rec.mean()))
