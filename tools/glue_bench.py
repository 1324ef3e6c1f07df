"""Stand-alone timing of the HBM-bound glue kernels at the benchmark shapes (B = TD_PROFILE_BATCH), with the
algorithmic bytes of each and the resulting fraction of the measured HBM peak."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinydiff import _lib as L, ops
dev = L.require_device("cuda:0")
lib = L.load()
B = int(os.environ.get("TD_PROFILE_BATCH", "128"))
peak = 6546.9
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def _graph_time(body, reps=5):
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        body()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        body()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def timeit(fn, n=8):
    """GPU time of one launch with a cold L2: graph of n x [flush 256 MiB, kernel] minus graph of n x [flush]."""
    def with_k():
        for _ in range(n):
            flush.zero_()
            fn()
    def without():
        for _ in range(n):
            flush.zero_()
    return (_graph_time(with_k) - _graph_time(without)) / n


def report(name, us, nbytes):
    print(f"{name:34s} {us:8.1f} us  {nbytes / 1e6:8.1f} MB  {nbytes / us / 1e3:8.1f} GB/s  {nbytes / us / 1e3 / peak:5.2f} of measured HBM peak", flush=True)


bf = torch.bfloat16
# final resize + conv
x = torch.randn(B, 32, 32, 64, device=dev).to(bf)
w = torch.randn(1, 3, 3, 64, device=dev)
b = torch.randn(1, device=dev)
yf = torch.empty(B, 1, 28, 28, device=dev)
report("final_resize_conv 32->28 C=64", timeit(lambda: L.check(lib.td_final_resize_conv(x.data_ptr(), L.TD_BF16, 64, 0, B, 32, 32, 64, w.data_ptr(),
       b.data_ptr(), 28, 28, yf.data_ptr(), L.stream_ptr()))), x.numel() * 2 + B * 784 * 4)
# upcat
for (hl, cu, hs, cs) in ((4, 512, 7, 512), (8, 256, 14, 256), (16, 128, 28, 128)):
    low = torch.randn(B, hl, hl, cu, device=dev).to(bf)
    skip = torch.randn(B, hs, hs, cs, device=dev).to(bf)
    temb = torch.randn(B, cs, device=dev)
    out = torch.empty(B, 2 * hl, 2 * hl, cu + cs, device=dev, dtype=bf)
    fn = lambda: L.check(lib.td_upcat_fwd(low.data_ptr(), skip.data_ptr(), temb.data_ptr(), cs, 0, out.data_ptr(), L.TD_BF16, B,
                                          2 * hl, 2 * hl, cu, hs, hs, cs, L.stream_ptr()))
    report(f"upcat {hl}->{2*hl} [{cu}|{cs}]", timeit(fn), (low.numel() + skip.numel() + out.numel()) * 2)
# maxpool
for (h, c) in ((28, 128), (14, 256), (7, 512)):
    xi = torch.randn(B, h, h, c, device=dev).to(bf)
    ho = (h + 1) // 2
    yo = torch.empty(B, ho, ho, c, device=dev, dtype=bf)
    fn = lambda: L.check(lib.td_maxpool2_fwd(xi.data_ptr(), yo.data_ptr(), L.TD_BF16, B, h, h, c, 1, L.stream_ptr()))
    report(f"maxpool {h}->{ho} C={c}", timeit(fn), (xi.numel() + yo.numel()) * 2)
# initial conv
xin = torch.randn(B, 1, 28, 28, device=dev)
w0 = ops.pack_conv_weight(torch.randn(64, 1, 3, 3, device=dev))
b0 = torch.randn(64, device=dev)
y0 = torch.empty(B, 28, 28, 64, device=dev, dtype=bf)
report("initial_conv 1->64 @28", timeit(lambda: ops.conv3x3(xin, w0, None, b0, False, 2, bf, x_nchw=True, out=y0)), B * 784 * 4 + B * 784 * 64 * 2)
# upcat backward (transposed resizes of both halves + the embedding-gradient reduction)
for (hl, cu, hs, cs) in ((4, 512, 7, 512), (8, 256, 14, 256), (16, 128, 28, 128)):
    dout = torch.randn(B, 2 * hl, 2 * hl, cu + cs, device=dev).to(bf)
    dlow = torch.empty(B, hl, hl, cu, device=dev, dtype=bf)
    dskip = torch.empty(B, hs, hs, cs, device=dev, dtype=bf)
    dtemb = torch.zeros(B, cs, device=dev)
    fn = lambda: L.check(lib.td_upcat_bwd(dout.data_ptr(), dlow.data_ptr(), dskip.data_ptr(), dtemb.data_ptr(), cs, 0, L.TD_BF16, B,
                                          2 * hl, 2 * hl, cu, hs, hs, cs, L.stream_ptr()))
    report(f"upcat_bwd {2*hl}->{hl} [{cu}|{cs}]", timeit(fn), (dout.numel() * 2 + dlow.numel() + dskip.numel()) * 2)
# maxpool backward
for (h, c) in ((28, 128), (14, 256), (7, 512)):
    xi = torch.randn(B, h, h, c, device=dev).to(bf)
    ho = (h + 1) // 2
    gy = torch.randn(B, ho, ho, c, device=dev).to(bf)
    gx = torch.empty_like(xi)
    fn = lambda: L.check(lib.td_maxpool2_bwd(xi.data_ptr(), gy.data_ptr(), gx.data_ptr(), L.TD_BF16, B, h, h, c, 1, 0, L.stream_ptr()))
    report(f"maxpool_bwd {ho}->{h} C={c}", timeit(fn), (2 * xi.numel() + gy.numel()) * 2)
# final resize 32 -> 28 forward / backward (train mode)
d1 = torch.randn(B, 32, 32, 64, device=dev).to(bf)
d1r = torch.empty(B, 28, 28, 64, device=dev, dtype=bf)
report("resize 32->28 C=64", timeit(lambda: L.check(lib.td_resize_bilinear_fwd(d1.data_ptr(), d1r.data_ptr(), L.TD_BF16, B, 32, 32, 28, 28, 64,
       L.stream_ptr()))), (d1.numel() + d1r.numel()) * 2)
report("resize_bwd 28->32 C=64", timeit(lambda: L.check(lib.td_resize_bilinear_bwd(d1r.data_ptr(), 64, 0, d1.data_ptr(), L.TD_BF16, B, 32, 32, 28, 28,
       64, L.stream_ptr()))), (d1.numel() + d1r.numel()) * 2)
