"""conv + BatchNorm + ReLU with and without the fused MaxPool2d(2) epilogue (conv_halo.cu): time per launch and the halo kernel's
wait-cycle counters, against the separate pooling kernel."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinydiff import _lib as L
dev = L.require_device("cuda:0")
lib = L.load()
B = int(os.environ.get("TD_PROFILE_BATCH", "128"))


def run(H, cin, cout, pool, iters=20):
    x = torch.randn(B, H, H, cin, device=dev).to(torch.bfloat16)
    w = torch.randn(cout, 3, 3, cin, device=dev).to(torch.bfloat16)
    y = torch.empty(B, H, H, cout, device=dev, dtype=torch.bfloat16)
    Hp = (H + 1) // 2
    py = torch.empty(B, Hp, Hp, cout, device=dev, dtype=torch.bfloat16)
    sc = torch.ones(cout, device=dev); sh = torch.zeros(cout, device=dev)
    d = L.ConvDesc()
    d.batch, d.height, d.width, d.cin, d.cout = B, H, H, cin, cout
    d.x_dtype = d.y_dtype = L.TD_BF16
    d.x, d.ldx, d.x_coff = x.data_ptr(), cin, 0
    d.y, d.ldy, d.y_coff = y.data_ptr(), cout, 0
    d.w, d.scale, d.shift, d.relu, d.stats, d.x_nchw, d.y_nchw = w.data_ptr(), sc.data_ptr(), sh.data_ptr(), 1, None, 0, 0
    if pool == "fused":
        d.pool_y, d.pool_ceil = py.data_ptr(), 1
    h = C.c_void_p()
    L.check(lib.td_conv3x3_plan_create(C.byref(h), C.byref(d), L.CONV_TC))
    fused = int(lib.td_conv3x3_pool_fused(h))
    st = L.stream_ptr()

    def step():
        lib.td_conv3x3_run(h, st)
        if pool == "separate":
            lib.td_maxpool2_fwd(y.data_ptr(), py.data_ptr(), L.TD_BF16, B, H, H, cout, 1, st)
    for _ in range(3):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    os.environ["TD_TC_HALO_DBG"] = "1"
    lib.td_conv3x3_run(h, st)
    buf = (C.c_ulonglong * (148 * 8))()
    L.check(lib.td_conv3x3_debug_counters(buf, 148 * 8))
    t = torch.tensor(list(buf), dtype=torch.float64).view(148, 8)
    cnt = t[t[:, 0] > 0].mean(0).tolist()
    os.environ.pop("TD_TC_HALO_DBG")
    lib.td_conv3x3_plan_destroy(h)
    return us, fused, cnt


print("kcyc: total | prod wait A, wait B | mma wait A, wait B, wait acc | epi wait acc, epi busy")
for H, ci, co in [(28, 128, 128), (14, 256, 256), (7, 512, 512)]:
    for pool in ("none", "separate", "fused"):
        us, fused, cnt = run(H, ci, co, pool)
        print(f"{H}x{H} {ci}->{co} pool={pool:9s} fused={fused} {us:7.1f} us  " + " ".join(f"{c / 1e3:7.1f}" for c in cnt), flush=True)
