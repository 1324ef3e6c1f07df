"""Debug aid (GPU box): train-mode forward/backward of one UNet against the CPU oracle, printing the
relative error of every intermediate activation, the loss and every parameter gradient."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib
import torch
from oracle import ddpm_oracle as O
from oracle.fixtures import init_state_dict, make_inputs
from tinydiff import _lib as L
from tinydiff.train import train_engine

name = sys.argv[1] if len(sys.argv) > 1 else "conditional_diffusion"
precision = sys.argv[2] if len(sys.argv) > 2 else "fp32"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 4
SPECS = {"diffusion": O.UNET_MNIST, "conditional_diffusion": O.UNET_COND, "conditional_diffusion_laion": O.UNET_LAION}
dev = L.require_device("cuda:0")
mod = importlib.import_module(f"tinydiff.{name}")
model = mod.NoiseModel()
sd = init_state_dict(name)
model.load_state_dict(sd, strict=True)
model.precision = precision
model = model.to(dev).train()
inp = make_inputs(name, B)
fp = mod.ForwardProcess()

def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))

leaf = {k: (v.detach().clone().requires_grad_(True) if O.is_param(k) else v.clone()) for k, v in sd.items()}
x_t = O.q_sample(fp.alphas_cumprod, inp["x0"], inp["t"], inp["noise"])
taps, stats = {}, {}
pred = O.unet_forward(SPECS[name], leaf, x_t, inp["t"], inp.get("cond"), training=True, new_stats=stats, taps=taps)
for v in taps.values():
    if v.requires_grad:
        v.retain_grad()
loss = torch.nn.functional.mse_loss(pred, inp["noise"])
loss.backward()

eng = train_engine(model, B, dev)
eng.refresh_weights()
eng.load_inputs(x_t.to(dev), inp["t"].to(dev), inp["cond"].to(dev) if "cond" in inp else None)
eng.launch_forward()
torch.cuda.synchronize()
print("== forward")
for k in ("x0", "e1", "e2", "e3", "b", "cat3", "d3", "cat2", "d2", "cat1", "d1", "d1r"):
    if k in eng.bufs and k in taps:
        print(f"  {k:6s} rel {rel(eng.bufs[k].float().permute(0, 3, 1, 2), taps[k].detach()):.3e}")
print(f"  eps    rel {rel(eng.eps, pred.detach()):.3e}")
n = eng.eps.numel()
d_eps = (2.0 / n) * (eng.eps - inp["noise"].to(dev))
eng.d_eps.copy_(d_eps)
eng.launch_backward()
torch.cuda.synchronize()
print("== activation gradients")
for k in ("d1r", "d1", "cat1", "d2", "cat2", "d3", "cat3", "b", "e3", "e2", "e1", "x0"):
    if k in eng.grads and k in taps and taps[k].grad is not None:
        print(f"  {k:6s} rel {rel(eng.grads[k].float().permute(0, 3, 1, 2), taps[k].grad):.3e}")
print("== per-layer dL/d(conv output) [backward order], as left in the shared dy buffer is not kept; re-run layer by layer")
st = L.stream_ptr()
eng.d_eps.copy_(d_eps)
for nm, fn in eng.bwd_ops:
    fn(st)
    if nm.startswith("bn:") and nm.endswith(":bwd"):
        layer = nm[3:-4]
        ref = taps.get(f"{layer}:y")
        if ref is not None and ref.grad is not None:
            Bq, Cq, Hq = ref.shape[0], ref.shape[1], ref.shape[2]
            mine = eng.dy[:ref.numel()].view(Bq, Hq, Hq, Cq).float().permute(0, 3, 1, 2)
            blk_out = {l[0]: l[4] for l in eng.layers}[layer]
            da_ref = taps[f"{layer}:a"].grad
            print(f"  {layer:14s} dy rel {rel(mine, ref.grad):.3e}   da rel {rel(eng.grads[blk_out].float().permute(0, 3, 1, 2), da_ref):.3e}"
                  f"   y rel {rel(eng.yraw[layer].permute(0, 3, 1, 2) + leaf[layer + '.bias'].detach().view(1, -1, 1, 1).to(dev), ref.detach()):.3e}")
torch.cuda.synchronize()
print("== parameter gradients")
for k, p in model.named_parameters():
    ref = leaf[k].grad
    if ref is None:
        print(f"  {k:34s} (no reference grad)")
        continue
    print(f"  {k:34s} rel {rel(eng.pgrad[k], ref):.3e}   |ref| {float(ref.norm()):.3e}")
