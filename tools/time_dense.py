"""Per-entry timing (CUDA events, eager) of the dense denoisers' train step at a large batch: which launches the 83 ms
(DiT) / 34 ms (latent MLP) at 65536 samples go to."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib
import torch
from tinydiff import _lib as L
from tinydiff.train import TrainStep

B = int(os.environ.get("TD_PROFILE_BATCH", "65536"))
dev = L.require_device("cuda:0")
for name in os.environ.get("TD_MODELS", "diffusion_transformer,latent_diffusion").split(","):
    mod = importlib.import_module(f"tinydiff.{name}")
    torch.manual_seed(0)
    kw = {"dropout": 0.0} if name == "diffusion_transformer" else {}
    model = mod.NoiseModel(**kw).to(dev).train()
    fp = mod.ForwardProcess()
    ts = TrainStep(model, fp, B, dev, use_graph=False)
    x0 = torch.randn(B, 20, device=dev)
    y = torch.randint(0, 10, (B,), device=dev)
    for _ in range(2):
        ts(x0, y)
    torch.cuda.synchronize()
    eng, st = ts.eng, L.stream_ptr()

    def time_ops(ops, iters=3):
        res = []
        for nm, fn in ops:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            fn(st); torch.cuda.synchronize()
            e0.record()
            for _ in range(iters): fn(st)
            e1.record(); torch.cuda.synchronize()
            res.append((nm, e0.elapsed_time(e1) / iters * 1e3))
        return res
    f, b = time_ops(eng.fwd_ops), time_ops(eng.bwd_ops)
    print(f"== {name} B={B}: forward {sum(v for _, v in f) / 1e3:.2f} ms, backward {sum(v for _, v in b) / 1e3:.2f} ms")
    for nm, v in sorted(f + b, key=lambda kv: -kv[1])[:24]:
        print(f"  {nm:34s} {v:9.1f} us")
    del ts, model
    torch.cuda.empty_cache()
