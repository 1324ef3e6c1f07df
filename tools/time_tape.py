"""Eval-forward time of the fused dense tape (latent MLP / DiT) at batch TD_PROFILE_BATCH (default 128): CUDA events around
`iters` back-to-back forwards."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib
import torch
from tinydiff import _lib as L
B = int(os.environ.get("TD_PROFILE_BATCH", "128"))
dev = L.require_device("cuda:0")
for name in ("latent_diffusion", "diffusion_transformer"):
    mod = importlib.import_module(f"tinydiff.{name}")
    torch.manual_seed(0)
    m = mod.NoiseModel().to(dev).eval()
    x = torch.randn(B, 20, device=dev)
    t = torch.randint(0, 1000, (B,), device=dev)
    y = torch.randint(0, 10, (B,), device=dev)
    with torch.no_grad():
        for _ in range(3):
            out = m(x, t, y)
        eng = m.engine(B, dev) if hasattr(m, "engine") else None
        torch.cuda.synchronize()
        st = L.stream_ptr()
        e = [v for v in getattr(m, "_engines", {}).values()][0] if eng is None else eng
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 200
        e0.record()
        for _ in range(iters):
            e._launch_tape(st)
        e1.record(); torch.cuda.synchronize()
    buf, n, nbar = e._tapes[False]
    print(f"{name}: B={B} tape ops {n}, barriers {nbar}: {e0.elapsed_time(e1) / iters * 1e3:.1f} us per forward")
