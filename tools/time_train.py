"""Per-kernel timing of one train step (CUDA events around every plan entry), B = TD_PROFILE_BATCH."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinydiff import _lib as L
from tinydiff.conditional_diffusion import ForwardProcess, NoiseModel
from tinydiff.train import TrainStep

B = int(os.environ.get("TD_PROFILE_BATCH", "128"))
prec = os.environ.get("TD_PRECISION", "bf16")
dev = L.require_device("cuda:0")
torch.manual_seed(0)
model = NoiseModel().to(dev).train()
model.precision = prec
fp = ForwardProcess()
ts = TrainStep(model, fp, B, dev, use_graph=False)
x0 = torch.rand(B, 1, 28, 28) * 2 - 1
y = torch.randint(0, 10, (B,))
for _ in range(3):
    ts(x0, y)
torch.cuda.synchronize()
eng = ts.eng
st = L.stream_ptr()
def time_ops(ops, iters=5):
    res = {}
    for name, fn in ops:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn(st); torch.cuda.synchronize()
        e0.record()
        for _ in range(iters): fn(st)
        e1.record(); torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / iters * 1e3
    return res
f = time_ops(eng.fwd_ops); b = time_ops(eng.bwd_ops)
print("forward  total %.1f us" % sum(f.values()))
for k, v in f.items(): print(f"  {k:22s} {v:8.1f}")
print("backward total %.1f us" % sum(b.values()))
for k, v in b.items(): print(f"  {k:22s} {v:8.1f}")
# whole step, eager and graph
def step_time(ts, n=10):
    for _ in range(3): ts.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): ts.run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("eager step ms", step_time(ts))
ts2 = TrainStep(model, fp, B, dev, use_graph=True)
ts2.load(x0, y)
print("graph step ms", step_time(ts2), " -> img/s", B / step_time(ts2) * 1e3)
print("conv flops/step %.1f GF" % (eng.conv_flops() / 1e9))
