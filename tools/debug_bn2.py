import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import ddpm_oracle as O
from oracle.fixtures import init_state_dict, make_inputs
from tinydiff import _lib as L, ops
from tinydiff.train import train_engine
import tinydiff.conditional_diffusion as mod
dev = L.require_device("cuda:0")
name="conditional_diffusion"; B=4
sd=init_state_dict(name); inp=make_inputs(name,B)
model=mod.NoiseModel(); model.load_state_dict(sd); model.precision="fp32"; model=model.to(dev).train()
_,_,ac=O.make_schedule()
x_t=O.q_sample(ac,inp["x0"],inp["t"],inp["noise"])
eng=train_engine(model,B,dev); eng.refresh_weights()
eng.load_inputs(x_t.to(dev), inp["t"].to(dev), inp["cond"].to(dev))
eng.launch_forward()
n=eng.eps.numel()
eng.d_eps.copy_((2.0/n)*(eng.eps-inp["noise"].to(dev)))
st=L.stream_ptr()
rel=lambda a,b: float((a.double().cpu()-b.double().cpu()).norm()/b.double().cpu().norm())
for nm,fn in eng.bwd_ops:
    fn(st)
    if nm=="bn:dec1.3:bwd":
        bn=eng.bn["dec1.3"]
        y=eng.yraw["dec1.3"]; da=eng.grads["d1"]
        dy2,dg2,db2=ops.bn_train_bwd(da,y,bn["scale"],bn["shift"],bn["mean"],bn["invstd"])
        mine=eng.dy[:y.numel()].view_as(y)
        print("engine dy vs ops dy", rel(mine,dy2), "dgamma", rel(eng.pgrad["dec1.4.weight"],dg2), "dbeta", rel(eng.pgrad["dec1.4.bias"],db2))
        print("rows fwd/bwd", bn["rows"], bn["rows_bwd"], "partials numel", eng.partials.numel())
        # recompute stats from y
        a2,sc2,sh2,mu2,iv2=ops.bn_train_fwd(y, model.dec1[4].weight, model.dec1[4].bias, None, None, None, None)
        print("scale", rel(bn["scale"],sc2), "shift", rel(bn["shift"],sh2), "mean", rel(bn["mean"],mu2), "invstd", rel(bn["invstd"],iv2))
        print("coef", bn["coef"][:, :4])
        break
