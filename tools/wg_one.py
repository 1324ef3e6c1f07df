"""One weight-gradient shape, a few launches (ncu target).  usage: wg_one.py H cin cout"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinydiff import _lib as L
dev = L.require_device("cuda:0")
lib = L.load()
H, cin, cout = (int(v) for v in sys.argv[1:4])
B = int(os.environ.get("TD_PROFILE_BATCH", "128"))
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
for _ in range(3):
    L.check(lib.td_conv3x3_wgrad_run(h, L.stream_ptr()))
torch.cuda.synchronize()
print("ok", float(dw.abs().mean()))
