"""Eval forward of the dense denoisers (latent MLP / DiT) at batch TD_PROFILE_BATCH (default 128): the cluster kernel
(csrc/dense_cluster.cu) against the grid-barrier tape (csrc/dense_fused.cu) and the one-launch-per-op path -- max abs
difference of eps, and CUDA-event time of `iters` back-to-back forwards plus of a graph-captured reverse step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib
import torch
from tinydiff import _lib as L
from tinydiff.dense import DenseEngine

B = int(os.environ.get("TD_PROFILE_BATCH", "128"))
dev = L.require_device("cuda:0")


def engine(m, mode):
    os.environ["TD_DENSE_FUSED"] = mode
    e = DenseEngine(m, B, dev, False, m.in_dim, m.emb_mode)
    m._declare(e)
    e.build()
    return e


for name in ("latent_diffusion", "diffusion_transformer"):
    mod = importlib.import_module(f"tinydiff.{name}")
    torch.manual_seed(0)
    m = mod.NoiseModel().to(dev).eval()
    with torch.no_grad():
        for p in m.parameters():
            p.add_(0.02 * torch.randn_like(p))
    x = torch.randn(B, 20, device=dev)
    t = torch.randint(0, 1000, (B,), device=dev)
    y = torch.randint(0, 10, (B,), device=dev)
    outs, st = {}, L.stream_ptr()
    for mode, label in (("0", "per-op"), ("2", "tape"), ("1", "cluster")):
        e = engine(m, mode)
        e.load_inputs(x, t, y)
        for _ in range(3):
            e.launch_forward()
        torch.cuda.synchronize()
        outs[label] = e.eps.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 200
        e0.record()
        for _ in range(iters):
            e.launch_forward()
        e1.record()
        torch.cuda.synchronize()
        kind = "cluster" if e._ctapes is not None else ("tape" if e._tapes is not None else "per-op")
        print(f"{name}: B={B} {label:8s} (runs as {kind}): {e0.elapsed_time(e1) / iters * 1e3:7.1f} us per forward", flush=True)
    ref = outs["per-op"]
    for label in ("tape", "cluster"):
        d = (outs[label] - ref).abs().max().item()
        print(f"{name}: {label} vs per-op max abs diff {d:.3e} (|eps| max {ref.abs().max().item():.3f})", flush=True)
    # graph-captured reverse steps through the public sampler
    for mode, label in (("2", "tape"), ("1", "cluster")):
        os.environ["TD_DENSE_FUSED"] = mode
        m._init_engines()
        fp = mod.ForwardProcess(num_timesteps=1000)
        mod.sample(None, m, fp, dev, n_samples=B, y=y, seed=1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        mod.sample(None, m, fp, dev, n_samples=B, y=y, seed=2)
        e1.record()
        torch.cuda.synchronize()
        print(f"{name}: B={B} {label:8s} 1000-step sampler {e0.elapsed_time(e1):.1f} ms = {e0.elapsed_time(e1):.1f} us per reverse step", flush=True)

if os.environ.get("TD_DENSE_CLUSTER_DBG"):
    import ctypes as C
    lib = L.load()
    os.environ["TD_DENSE_FUSED"] = "1"
    for name in ("latent_diffusion", "diffusion_transformer"):
        mod = importlib.import_module(f"tinydiff.{name}")
        m = mod.NoiseModel().to(dev).eval()
        e = DenseEngine(m, B, dev, False, m.in_dim, m.emb_mode)
        m._declare(e)
        e.build()
        e.load_inputs(torch.randn(B, 20, device=dev), torch.randint(0, 1000, (B,), device=dev), torch.randint(0, 10, (B,), device=dev))
        for _ in range(3):
            e.launch_forward()
        ncta = 8 * min(-(-B // e._ctape_rows), 15)
        n = ncta * 16
        buf = (C.c_ulonglong * n)()
        L.check(lib.td_dense_cluster_debug_counters(buf, n), "dbg")
        v = torch.tensor(list(buf), dtype=torch.float64).view(ncta, 16)
        t0 = v[:, 0].min()
        print(f"{name}: CTA start (us after first) min/max {((v[:,0]-t0)/1e3).min():.1f}/{((v[:,0]-t0)/1e3).max():.1f}; end min/max "
              f"{((v[:,1]-t0)/1e3).min():.1f}/{((v[:,1]-t0)/1e3).max():.1f}")
        per_cluster = ((v[:, 0] - t0) / 1e3).view(-1, 8)[:, 0]
        print("  cluster start times (us):", [round(float(x), 1) for x in per_cluster])
        for i, lab in enumerate(("wait", "fma", "epi:push", "barrier", "rowops", "prefetch", "epi:sts+sync", "epi:reduce+math")):
            c = v[:, 2 + i]
            print(f"  {lab:9s} kcycles mean {c.mean()/1e3:8.1f} min {c.min()/1e3:8.1f} max {c.max()/1e3:8.1f}")
