import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from oracle import ddpm_oracle as O
from oracle.fixtures import init_state_dict, make_inputs
from tinydiff import _lib as L
import tinydiff.diffusion_transformer as mod
dev = L.require_device("cuda:0")
name = "diffusion_transformer"; B = 8
sd = init_state_dict(name, perturb=False)
model = mod.NoiseModel(dropout=0.0); model.load_state_dict(sd); model = model.to(dev).train()
inp = make_inputs(name, B)
_, _, ac = O.make_schedule()
x_t = O.q_sample(ac, inp["x0"], inp["t"], inp["noise"])
leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
T = {}
def keep(n, v):
    v.retain_grad(); T[n] = v; return v
D = 256
tt = (inp["t"] / 1000).unsqueeze(-1).float()
h = keep("h", F.silu(F.linear(tt, leaf["time_embedding.0.weight"], leaf["time_embedding.0.bias"])))
emb = keep("emb", F.linear(h, leaf["time_embedding.2.weight"], leaf["time_embedding.2.bias"]) + F.embedding(inp["cond"], leaf["class_embedding.weight"]))
cur = keep("x.0", F.linear(x_t, leaf["input_proj.weight"], leaf["input_proj.bias"]) + emb + leaf["pos_encoding"].view(1, -1))
for i in range(4):
    p = f"transformer_blocks.{i}."
    v = keep(f"blk{i}.v", F.linear(cur, leaf[p + "attention.in_proj_weight"][2 * D:], leaf[p + "attention.in_proj_bias"][2 * D:]))
    s1 = keep(f"blk{i}.s1", cur + F.linear(v, leaf[p + "attention.out_proj.weight"], leaf[p + "attention.out_proj.bias"]))
    x1 = keep(f"blk{i}.x1", F.layer_norm(s1, (D,), leaf[p + "norm1.weight"], leaf[p + "norm1.bias"]))
    u = keep(f"blk{i}.u", F.gelu(F.linear(x1, leaf[p + "ff.0.weight"], leaf[p + "ff.0.bias"])))
    s2 = keep(f"blk{i}.s2", x1 + F.linear(u, leaf[p + "ff.2.weight"], leaf[p + "ff.2.bias"]))
    cur = keep(f"blk{i}.x2", F.layer_norm(s2, (D,), leaf[p + "norm2.weight"], leaf[p + "norm2.bias"]))
xf = keep("xf", F.layer_norm(cur, (D,), leaf["final_layer.0.weight"], leaf["final_layer.0.bias"]))
pred = F.linear(xf, leaf["final_layer.1.weight"], leaf["final_layer.1.bias"])
loss = F.mse_loss(pred, inp["noise"]); loss.backward()
eng = model.engine(B, dev, training=True)
eng.load_inputs(x_t.to(dev), inp["t"].to(dev), inp["cond"].to(dev))
eng.launch_forward()
rel = lambda a, b: float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm().clamp_min(1e-30))
print("eps", rel(eng.eps, pred.detach()))
eng.d_eps.copy_((2.0 / pred.numel()) * (eng.eps - inp["noise"].to(dev)))
st = L.stream_ptr()
for nm, fn in eng.bwd_ops:
    fn(st)
torch.cuda.synchronize()
for n, v in T.items():
    print(f"{n:12s} act {rel(eng.bufs[n], v.detach()):.2e}  grad {rel(eng.gbufs[n], v.grad):.2e}")
for k in ("final_layer.1.weight", "final_layer.1.bias", "final_layer.0.weight", "transformer_blocks.3.ff.2.weight", "transformer_blocks.3.ff.2.bias",
          "transformer_blocks.3.norm2.weight", "transformer_blocks.3.ff.0.weight", "transformer_blocks.3.attention.out_proj.weight", "input_proj.weight", "pos_encoding"):
    print(f"{k:48s} {rel(eng.pgrad[k], leaf[k].grad):.2e}")
print([n for n, _ in eng.bwd_ops][:12])
