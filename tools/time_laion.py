"""Per-entry timing of one eval forward of the LAION latent UNet (config 5) at batch TD_PROFILE_BATCH (default 256)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinydiff import _lib as L
from tinydiff.conditional_diffusion_laion import NoiseModel
B = int(os.environ.get("TD_PROFILE_BATCH", "256"))
dev = L.require_device("cuda:0")
torch.manual_seed(0)
m = NoiseModel().to(dev).eval()
eng = m.engine(B, dev)
eng.refresh_weights()
eng.x_in.normal_(); eng.text_in.normal_(); eng.t_in.fill_(500)
eng.launch(); torch.cuda.synchronize()
st = L.stream_ptr()
tot = 0.0
for n, fn in eng.ops:
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(st); torch.cuda.synchronize()
    a.record()
    for _ in range(4): fn(st)
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / 4 * 1e3
    tot += us
    fl = eng.plans[n].flops if n in eng.plans else 0.0
    print(f"{n:22s} {us:8.1f} us  {fl / 1e9:7.2f} GFLOP  {fl / (us * 1e-6) / 1e12 if us else 0:7.1f} TFLOP/s  engine={eng.engines.get(n)}")
print("total", tot)
