"""Sweep the tcgen05 conv engine's tuning knobs (pipeline stages -> CTAs/SM, BLOCK_N) per UNet layer shape."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinydiff import ops, _lib as L
dev = L.require_device("cuda:0")
lib = L.load()
def run(B, H, cin, cout, env, iters=20):
    for k in ("TD_TC_STAGES", "TD_TC_BLOCK_N", "TD_TC_SPLIT_K"):
        os.environ.pop(k, None)
    os.environ.update(env)
    x = torch.randn(B, H, H, cin, device=dev).to(torch.bfloat16)
    w = torch.randn(cout, 3, 3, cin, device=dev).to(torch.bfloat16)
    y = torch.empty(B, H, H, cout, device=dev, dtype=torch.bfloat16)
    d = L.ConvDesc()
    d.batch, d.height, d.width, d.cin, d.cout = B, H, H, cin, cout
    d.x_dtype = d.y_dtype = L.TD_BF16
    d.x, d.ldx, d.x_coff = x.data_ptr(), cin, 0
    d.y, d.ldy, d.y_coff = y.data_ptr(), cout, 0
    d.w, d.scale, d.shift, d.relu, d.stats, d.x_nchw, d.y_nchw = w.data_ptr(), None, None, 0, None, 0, 0
    need = int(lib.td_conv3x3_splitk_workspace(C.byref(d)))
    ws = torch.empty(max(need, 1), device=dev)
    d.splitk_ws = ws.data_ptr() if need > 0 else None
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
    lib.td_conv3x3_plan_destroy(h)
    return ms * 1e3, fl / ms / 1e9
shapes = [(28, 64, 128), (28, 128, 128), (14, 128, 256), (14, 256, 256), (7, 256, 512), (7, 512, 512), (4, 512, 512),
          (8, 1024, 256), (8, 256, 256), (16, 512, 128), (16, 128, 128), (32, 256, 64), (32, 64, 64)]
B = int(os.environ.get("TD_PROFILE_BATCH", "128"))
cfgs = [{}, {"TD_TC_SPLIT_K": "1"}, {"TD_TC_SPLIT_K": "2"}, {"TD_TC_SPLIT_K": "3"}, {"TD_TC_SPLIT_K": "4"},
        {"TD_TC_BLOCK_N": "128"}, {"TD_TC_BLOCK_N": "128", "TD_TC_SPLIT_K": "2"}]
print("shape".ljust(18) + "".join((",".join(f"{k[6:]}={v}" for k, v in c.items()) or "default").rjust(22) for c in cfgs))
tot = [0.0] * len(cfgs)
for H, ci, co in shapes:
    row = f"{H}x{H} {ci}->{co}".ljust(18)
    for i, c in enumerate(cfgs):
        try:
            us, tf = run(B, H, ci, co, c)
            row += f"{us:9.1f}us {tf:6.0f}TF".rjust(22)
            tot[i] += us
        except Exception as e:
            row += "n/a".rjust(22); tot[i] += 1e9
    print(row, flush=True)
print("total us".ljust(18) + "".join(f"{t:22.1f}" for t in tot))
