"""Diagnostics for the tcgen05 conv engine on a B200: per-shape error + timing (CUDA events)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from tinydiff import ops, _lib as L

dev = L.require_device("cuda:0")

def nhwc(x): return x.permute(0, 2, 3, 1).contiguous()
def nchw(x): return x.permute(0, 3, 1, 2).contiguous()

def case(B, H, cin, cout, check=True, iters=20):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, cin, H, H, generator=g).to(torch.bfloat16)
    w = (torch.randn(cout, cin, 3, 3, generator=g) / (9 * cin) ** 0.5)
    xd = nhwc(x).to(dev)
    wp = ops.pack_conv_weight(w.to(dev), torch.bfloat16)
    out = torch.empty(B, H, H, cout, device=dev, dtype=torch.bfloat16)
    ops.conv3x3(xd, wp, None, None, False, L.CONV_TC, torch.bfloat16, out=out)
    torch.cuda.synchronize()
    msg = f"B={B} H={H} cin={cin} cout={cout}"
    if check:
        want = F.conv2d(x.float().to(dev), w.to(torch.bfloat16).float().to(dev), padding=1)
        got = nchw(out).float()
        err = (got - want).norm() / want.norm()
        msg += f" rel={float(err):.3e}"
        if err > 1e-2:
            d = (got - want).abs()
            bad = (d > 0.05).float()
            msg += f" bad_frac={float(bad.mean()):.3f} by_n={bad.mean((1,2,3)).tolist()[:4]} by_h={bad.mean((0,1,3)).tolist()[:8]} by_w={bad.mean((0,1,2)).tolist()[:8]} by_c64={[float(bad[:, i:i+64].mean()) for i in range(0, cout, 64)][:8]}"
    # timing via plan reuse
    import ctypes as C
    lib = L.load()
    d = L.ConvDesc()
    d.batch, d.height, d.width, d.cin, d.cout = B, H, H, cin, cout
    d.x_dtype = d.y_dtype = L.TD_BF16
    d.x, d.ldx, d.x_coff = xd.data_ptr(), cin, 0
    d.y, d.ldy, d.y_coff = out.data_ptr(), cout, 0
    d.w, d.scale, d.shift, d.relu, d.stats, d.x_nchw, d.y_nchw = wp.data_ptr(), None, None, 0, None, 0, 0
    h = C.c_void_p()
    L.check(lib.td_conv3x3_plan_create(C.byref(h), C.byref(d), L.CONV_TC))
    st = L.stream_ptr()
    for _ in range(3): lib.td_conv3x3_run(h, st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): lib.td_conv3x3_run(h, st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = lib.td_conv3x3_flops(h)
    msg += f" time={ms*1e3:.1f}us TFLOPs={fl/ms/1e9:.1f}"
    lib.td_conv3x3_plan_destroy(h)
    print(msg, flush=True)

if __name__ == "__main__":
    small = [(2, 8, 64, 64), (2, 8, 128, 128), (3, 28, 64, 128), (3, 7, 256, 512), (3, 14, 128, 256), (3, 4, 512, 512)]
    for c in small: case(*c)
    full = [(28, 64, 128), (28, 128, 128), (14, 128, 256), (14, 256, 256), (7, 256, 512), (7, 512, 512), (4, 512, 512),
            (8, 1024, 256), (8, 256, 256), (16, 512, 128), (16, 128, 128), (32, 256, 64), (32, 64, 64)]
    for H, ci, co in full: case(128, H, ci, co, check=(ci <= 256))
