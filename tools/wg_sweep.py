"""Sweep the tcgen05 weight-gradient engine's knobs per UNet layer shape (B = TD_PROFILE_BATCH)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinydiff import _lib as L
dev = L.require_device("cuda:0")
lib = L.load()
def run(B, H, cin, cout, env, iters=10):
    for k in ("TD_WG_STAGES", "TD_WG_BLOCK_N", "TD_WG_TARGET_CTAS", "TD_WG_HALO"):
        os.environ.pop(k, None)
    os.environ.update(env)
    x = torch.randn(B, H, H, cin, device=dev).to(torch.bfloat16)
    dy = torch.randn(B, H, H, cout, device=dev).to(torch.bfloat16)
    dw = torch.empty(cout, cin, 3, 3, device=dev)
    d = L.WgradDesc()
    d.batch, d.height, d.width, d.cin, d.cout = B, H, H, cin, cout
    d.x_dtype = d.dy_dtype = L.TD_BF16
    d.x, d.ldx, d.x_coff, d.x_nchw = x.data_ptr(), cin, 0, 0
    d.dy, d.lddy, d.dy_coff, d.dy_nchw = dy.data_ptr(), cout, 0, 0
    d.dw = dw.data_ptr()
    ws = torch.empty(max(int(lib.td_conv3x3_wgrad_workspace(C.byref(d), L.CONV_TC)), 1), device=dev)
    d.workspace = ws.data_ptr()
    h = C.c_void_p()
    L.check(lib.td_conv3x3_wgrad_plan_create(C.byref(h), C.byref(d), L.CONV_TC))
    st = L.stream_ptr()
    for _ in range(3): lib.td_conv3x3_wgrad_run(h, st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): lib.td_conv3x3_wgrad_run(h, st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    lib.td_conv3x3_wgrad_plan_destroy(h)
    fl = 2.0 * B * H * H * cout * 9 * cin
    return ms * 1e3, fl / ms / 1e9, ws.numel() * 4 / 2**20
shapes = [(28, 64, 128), (28, 128, 128), (14, 128, 256), (14, 256, 256), (7, 256, 512), (7, 512, 512), (4, 512, 512),
          (8, 1024, 256), (8, 256, 256), (16, 512, 128), (16, 128, 128), (32, 256, 64), (32, 64, 64)]
B = int(os.environ.get("TD_PROFILE_BATCH", "128"))
cfgs = [{"TD_WG_HALO": "0"}, {}, {"TD_WG_STAGES": "2"}, {"TD_WG_TARGET_CTAS": "74"}, {"TD_WG_TARGET_CTAS": "296"}]
print("shape".ljust(18) + "".join((",".join(f"{k[6:]}={v}" for k, v in c.items()) or "default").rjust(24) for c in cfgs))
tot = [0.0] * len(cfgs)
for H, ci, co in shapes:
    row = f"{H}x{H} {ci}->{co}".ljust(18)
    for i, c in enumerate(cfgs):
        try:
            us, tf, mb = run(B, H, ci, co, c)
            row += f"{us:7.1f}us {tf:5.0f}TF {mb:4.0f}MB".rjust(24)
            tot[i] += us
        except Exception as e:
            row += "n/a".rjust(24); tot[i] += 1e9
    print(row, flush=True)
print("total us".ljust(18) + "".join(f"{t:24.1f}" for t in tot))
