"""Per-layer timing of the tcgen05 conv engine: per-tap kernel vs halo kernel (layouts, ring depths), plus the
halo kernel's per-role wait-cycle counters (TD_TC_HALO_DBG)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinydiff import _lib as L
dev = L.require_device("cuda:0")
lib = L.load()
KEYS = ("TD_TC_HALO", "TD_TC_HALO_MODE", "TD_TC_HALO_NA", "TD_TC_HALO_NB", "TD_TC_HALO_DBG", "TD_TC_HALO_ROT", "TD_TC_STAGES", "TD_TC_BLOCK_N",
        "TD_TC_SPLIT_K")


def run(B, H, cin, cout, env, iters=20, dbg=False):
    for k in KEYS:
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
    for _ in range(3):
        lib.td_conv3x3_run(h, st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        lib.td_conv3x3_run(h, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = lib.td_conv3x3_flops(h)
    cnt = None
    if dbg:
        os.environ["TD_TC_HALO_DBG"] = "1"
        lib.td_conv3x3_run(h, st)
        buf = (C.c_ulonglong * (148 * 8))()
        L.check(lib.td_conv3x3_debug_counters(buf, 148 * 8))
        t = torch.tensor(list(buf), dtype=torch.float64).view(148, 8)
        cnt = t[t[:, 0] > 0].mean(0).tolist()
        os.environ.pop("TD_TC_HALO_DBG")
    lib.td_conv3x3_plan_destroy(h)
    return ms * 1e3, fl / ms / 1e9, cnt


shapes = [(28, 64, 128), (28, 128, 128), (14, 128, 256), (14, 256, 256), (7, 256, 512), (7, 512, 512),
          (16, 512, 128), (16, 128, 128), (32, 256, 64), (32, 64, 64)]
B = int(os.environ.get("TD_PROFILE_BATCH", "128"))
cfgs = [("per-tap", {"TD_TC_HALO": "0"}), ("halo", {}), ("noA,noB", {"TD_TC_HALO_DBG": "6"}), ("flat", {"TD_TC_HALO_MODE": "1"}),
        ("strip", {"TD_TC_HALO_MODE": "2"}), ("dx", {"TD_TC_HALO_MODE": "3"})]
print("shape".ljust(18) + "".join(n.rjust(20) for n, _ in cfgs))
tot = [0.0] * len(cfgs)
for H, ci, co in shapes:
    row = f"{H}x{H} {ci}->{co}".ljust(18)
    for i, (_, c) in enumerate(cfgs):
        try:
            us, tf, _ = run(B, H, ci, co, c)
            row += f"{us:9.1f}us {tf:5.0f}TF".rjust(20)
            tot[i] += us
        except Exception as e:
            row += "n/a".rjust(20)
            tot[i] += 1e9
    print(row, flush=True)
print("total us".ljust(18) + "".join(f"{t:20.1f}" for t in tot))
print("\nhalo wait-cycle counters, mean over CTAs (kcyc): total | prod wait A, wait B | mma wait A, wait B, wait acc | epi wait acc, epi busy")
for H, ci, co in shapes:
    us, tf, cnt = run(B, H, ci, co, {}, dbg=True)
    print(f"{H}x{H} {ci}->{co}".ljust(18) + f"{us:8.1f}us " + " ".join(f"{c / 1e3:8.1f}" for c in cnt), flush=True)
