"""Throughput of the other BASELINE.json configurations (3: DiT, 4: latent MLP, 5: LAION latent UNet) at the batch sweeps of
SURVEY.md 8d: fused train step (one CUDA graph) and reverse step of the sampler (graph-captured, T = 100 here; a 1000-step
sample costs 10x).  Not bench lines -- these configurations are parity-test cases -- but measured evidence for DESIGN.md."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinydiff import _lib as L
from tinydiff.train import TrainStep

dev = L.require_device("cuda:0")
T = 100


def ev_time(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run(name, B):
    import importlib
    mod = importlib.import_module(f"tinydiff.{name}")
    torch.manual_seed(0)
    kw = {"dropout": 0.0} if name == "diffusion_transformer" else {}
    model = mod.NoiseModel(**kw).to(dev)
    fp = mod.ForwardProcess(num_timesteps=T)
    g = torch.Generator().manual_seed(1)
    if name == "conditional_diffusion_laion":
        x0 = 0.18215 * torch.randn(B, 4, 32, 32, generator=g)
        cond = torch.randn(B, 768, generator=g)
    else:
        x0 = torch.randn(B, 20, generator=g)
        cond = torch.randint(0, 10, (B,), generator=g)
    # train
    model.train()
    ts = TrainStep(model, fp, B, dev, use_graph=True)
    ts.load(x0, cond)
    for _ in range(3):
        ts.run()
    ms_train = ev_time(ts.run, 20)
    ts.close()
    # sampler
    model.eval()
    if name == "conditional_diffusion_laion":
        f = lambda: mod.sample(model, fp, dev, text_embeds=cond.to(dev), seed=3)
    else:
        f = lambda: mod.sample(None, model, fp, dev, n_samples=B, y=cond.to(dev), seed=3)
    f()
    ms_sample = ev_time(f, 3)
    print(f"{name:30s} B={B:6d}  train {ms_train:8.3f} ms/step {B / ms_train * 1e3:12.0f} samples/s   "
          f"reverse step {ms_sample / T * 1e3:8.1f} us -> 1000-step sampler {B / (ms_sample / T * 1000) * 1e3:10.1f} samples/s", flush=True)
    del ts, model
    torch.cuda.empty_cache()


for name, batches in (("conditional_diffusion_laion", (8, 64, 256)), ("latent_diffusion", (128, 4096, 65536)),
                      ("diffusion_transformer", (128, 4096, 65536))):
    for B in batches:
        try:
            run(name, B)
        except Exception as e:          # keep the sweep going; the failure is part of the record
            print(f"{name:30s} B={B:6d}  FAILED: {type(e).__name__}: {str(e)[:200]}", flush=True)
