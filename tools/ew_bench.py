"""HBM roofline of the elementwise DDPM kernels at a working set far above the 126 MB L2 (at the benchmark
batch these tensors are 0.4 MB, L2-resident and launch-bound, so the bandwidth fraction is quoted here at 64 M elements
= 256 MB per tensor and the configuration-size latency beside it).  CUDA events, 10 launches after 3 warm-ups."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinydiff import _lib as L
from tinydiff.conditional_diffusion import ForwardProcess

dev = L.require_device("cuda:0")
lib = L.load()
peak = 6546.9
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
fp = ForwardProcess()
tab = fp._tables(dev)
st = L.stream_ptr()


ONCE = os.environ.get("TD_EW_ONCE") == "1"        # one launch per kernel and size: the ncu pass


def timeit(fn, reps=10):
    if ONCE:
        fn()
        torch.cuda.synchronize()
        return float("nan")
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3          # us


def report(name, us, nbytes):
    print(f"{name:44s} {us:9.1f} us {nbytes / 1e6:9.1f} MB {nbytes / us / 1e3:8.1f} GB/s  {nbytes / us / 1e3 / peak:5.2f} of measured HBM peak "
          f"({peak:.0f} GB/s)", flush=True)


for B, per, tag in ((16384, 4096, "64 M elements"), (128, 784, "bench size, 128 x 784")):
    n = B * per
    x0 = torch.rand(n, device=dev) * 2 - 1
    noise = torch.randn(n, device=dev)
    x_t = torch.empty(n, device=dev)
    t = torch.randint(0, 1000, (B,), device=dev)
    report(f"qsample (12 B/elem)            [{tag}]", timeit(lambda: L.check(lib.td_qsample(
        x0.data_ptr(), noise.data_ptr(), t.data_ptr(), tab["abar"].data_ptr(), x_t.data_ptr(), B, per, 1000, None, st))), 12 * n)
    grad = torch.empty(n, device=dev)
    loss = torch.zeros(1, device=dev)
    partials = torch.zeros(int(lib.td_mse_num_partials(n)), device=dev)
    counter = torch.zeros(1, device=dev, dtype=torch.int32)
    report(f"mse_grad (12 B/elem)           [{tag}]", timeit(lambda: L.check(lib.td_mse_grad(
        x_t.data_ptr(), noise.data_ptr(), grad.data_ptr(), loss.data_ptr(), partials.data_ptr(), counter.data_ptr(), n, 1.0 / n, st))),
        12 * n)
    t_dev = torch.tensor([500], dtype=torch.int32, device=dev)
    report(f"psample_step, injected z (16 B/elem) [{tag}]", timeit(lambda: L.check(lib.td_psample_step(
        x_t.data_ptr(), grad.data_ptr(), noise.data_ptr(), 0, tab["coef"].data_ptr(), t_dev.data_ptr(), n, 1000, None, st))), 16 * n)
    seed = torch.tensor([7, 0], dtype=torch.int64, device=dev)
    report(f"psample_step, Philox z (12 B/elem)   [{tag}]", timeit(lambda: L.check(lib.td_psample_step(
        x_t.data_ptr(), grad.data_ptr(), None, 0, tab["coef"].data_ptr(), t_dev.data_ptr(), n, 1000, seed.data_ptr(), st))), 12 * n)

# fused Adam: one 64 M-element tensor (28 B/param) and the UNet's 11.18 M parameters as 90 tensors' worth of chunks
for n, tag in ((64 << 20, "64 M parameters"), (11182273 // 4 * 4, "11.18 M parameters (UNet)")):
    p_, g_, m_, v_ = (torch.randn(n, device=dev) * 0.01 for _ in range(4))
    v_.abs_()
    CH = 16384
    chunks = (n + CH - 1) // CH
    t64 = lambda vals: torch.tensor(vals, dtype=torch.int64, device=dev)
    tp, tg, tm, tv = t64([p_.data_ptr()]), t64([g_.data_ptr()]), t64([m_.data_ptr()]), t64([v_.data_ptr()])
    numel = t64([n])
    ct = torch.zeros(chunks, dtype=torch.int32, device=dev)
    co = torch.arange(chunks, dtype=torch.int64, device=dev) * CH
    step = torch.ones(1, dtype=torch.int32, device=dev)
    gs = torch.ones(1, device=dev)
    report(f"adam_multi (28 B/param)        [{tag}]", timeit(lambda: L.check(lib.td_adam_multi(
        tp.data_ptr(), tg.data_ptr(), tm.data_ptr(), tv.data_ptr(), numel.data_ptr(), ct.data_ptr(), co.data_ptr(), chunks, CH,
        step.data_ptr(), 1e-3, None, 0.9, 0.999, 1e-8, gs.data_ptr(), None, st))), 28 * n)
