import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import ddpm_oracle as O
from oracle.fixtures import init_state_dict, make_inputs
from tinydiff import _lib as L, ops
dev = L.require_device("cuda:0")
name="conditional_diffusion"; B=4
sd=init_state_dict(name); inp=make_inputs(name,B)
_,_,ac=O.make_schedule()
leaf={k:(v.clone().requires_grad_(True) if O.is_param(k) else v) for k,v in sd.items()}
x_t=O.q_sample(ac,inp["x0"],inp["t"],inp["noise"])
taps={}
pred=O.unet_forward(O.UNET_COND,leaf,x_t,inp["t"],inp["cond"],training=True,new_stats={},taps=taps)
for v in taps.values():
    if v.requires_grad: v.retain_grad()
torch.nn.functional.mse_loss(pred,inp["noise"]).backward()
rel=lambda a,b: float((a.double().cpu()-b.double().cpu()).norm()/b.double().cpu().norm())
for layer, bnk in (("dec1.3","dec1.4"),("dec1.0","dec1.1"),("enc1.0","enc1.1")):
    y=taps[layer+":y"].detach(); da=taps[layer+":a"].grad; dy_ref=taps[layer+":y"].grad
    bias=sd[layer+".bias"]
    yd=(y-bias.view(1,-1,1,1)).permute(0,2,3,1).contiguous().to(dev)
    a,scale,shift,mean,invstd=ops.bn_train_fwd(yd, sd[bnk+".weight"].to(dev), sd[bnk+".bias"].to(dev), bias.to(dev), None, None, None)
    print(layer, "a rel", rel(a.permute(0,3,1,2), taps[layer+":a"].detach()))
    dy,dg,db=ops.bn_train_bwd(da.permute(0,2,3,1).contiguous().to(dev), yd, scale, shift, mean, invstd)
    err=(dy.permute(0,3,1,2).cpu()-dy_ref)
    print(layer, "dy rel", rel(dy.permute(0,3,1,2), dy_ref), "dgamma", rel(dg, leaf[bnk+".weight"].grad), "dbeta", rel(db, leaf[bnk+".bias"].grad))
    pc=(err.pow(2).sum((0,2,3)).sqrt()/dy_ref.pow(2).sum((0,2,3)).sqrt())
    print("   per-channel err (first 16):", [f"{float(v):.1e}" for v in pc[:16]])
    print("   |mean|/std per channel (first 8):", [f"{float(v):.1f}" for v in (mean.cpu().abs()*invstd.cpu())[:8]])
