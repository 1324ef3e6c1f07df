"""Training path of the conv-UNet denoisers: train-mode forward (batch-statistics BatchNorm),
the full backward, and the fused train step (diffusion.py:220-236).

``UNetTrainEngine`` is the train-mode sibling of ``unet.UNetEngine``: a fixed launch sequence of
libtinydiff kernels over preallocated NHWC buffers, for one (batch, precision) pair.  Two ways in:

* ``unet_train_forward`` -- what ``NoiseModel.forward`` calls in train mode: a
  ``torch.autograd.Function`` whose backward runs the backward plan, so the reference's own five
  statements (``q_sample; model(x_t, t); F.mse_loss; loss.backward(); optimizer.step()``) work
  unchanged and ``.grad`` of every parameter is populated.
* ``TrainStep`` -- the same work without autograd: q_sample -> forward -> MSE(+grad) -> backward ->
  (gradient all-reduce) -> fused Adam, captured in one CUDA graph.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, Dict, List, Optional, Tuple

import torch

from . import _lib as L
from .dp import BucketReducer, grad_ready_index, plan_buckets
from .unet import UNetConfig, _CONVS, _ConvPlan, _pool, alloc_splitk_ws

BN_EPS_DEFAULT = 1e-5


class _WgradPlan:
    def __init__(self, desc: L.WgradDesc, engine: int):
        self.lib = L.load()
        self.handle = C.c_void_p()
        self.desc = desc
        L.check(self.lib.td_conv3x3_wgrad_plan_create(C.byref(self.handle), C.byref(desc), engine),
                "td_conv3x3_wgrad_plan_create")

    def run(self, stream: int) -> None:
        L.check(self.lib.td_conv3x3_wgrad_run(self.handle, stream), "td_conv3x3_wgrad_run")

    def __del__(self):
        try:
            if self.handle:
                self.lib.td_conv3x3_wgrad_plan_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass


class UNetTrainEngine:
    """Train-mode forward + backward plan.  ``precision``: "bf16" (tcgen05 convolutions, bf16 NHWC
    activations and activation gradients, fp32 statistics / parameter gradients) or "fp32"
    (FFMA convolutions; the tight-tolerance parity path)."""

    def __init__(self, cfg: UNetConfig, module: torch.nn.Module, batch: int, device: torch.device,
                 precision: str = "bf16"):
        assert precision in ("bf16", "fp32")
        self.cfg, self.module, self.B, self.device, self.precision = cfg, module, batch, device, precision
        self.lib = L.load()
        self.act = torch.bfloat16 if precision == "bf16" else torch.float32
        self.adt = L.dtype_code(self.act)
        c0, c1, c2, c3 = cfg.enc
        d3, d2, d1 = cfg.dec
        s0 = cfg.image
        s1 = _pool(s0, cfg.ceil_pool)
        s2 = _pool(s1, cfg.ceil_pool)
        s3 = _pool(s2, cfg.ceil_pool)
        u3, u2, u1 = 2 * s3, 4 * s3, 8 * s3
        self.S = dict(s0=s0, s1=s1, s2=s2, s3=s3, u3=u3, u2=u2, u1=u1)
        B = batch
        dev = device
        # initial_conv's output (and its gradient) is allocated 64 channels wide when the model has fewer (the LAION latent
        # UNet: 32): the upper channels stay zero and enc1.0's packed operands are zero-padded to match, so its forward,
        # data gradient and weight gradient run on the tcgen05 engine like every other layer (see unet.UNetEngine.c0p)
        self.c0p = c0
        if precision == "bf16" and c0 % 64 != 0 and c1 % 64 == 0 and c0 % 8 == 0:
            self.c0p = (c0 + 63) // 64 * 64
        c0p = self.c0p

        def buf(h, c, dtype=None):
            return torch.zeros(B, h, h, c, device=dev, dtype=dtype or self.act)

        # (block name) -> input buffer, cin, spatial size, output buffer name
        self.layers: List[Tuple[str, str, int, int, str, int]] = [
            # name, input buffer, cin, size, output buffer, cout
            ("enc1.0", "x0", c0p, s0, "enc1a", c1), ("enc1.3", "enc1a", c1, s0, "e1", c1),
            ("enc2.0", "p1", c1, s1, "enc2a", c2), ("enc2.3", "enc2a", c2, s1, "e2", c2),
            ("enc3.0", "p2", c2, s2, "enc3a", c3), ("enc3.3", "enc3a", c3, s2, "e3", c3),
            ("bottleneck.0", "p3", c3, s3, "b", cfg.bott),
            ("dec3.0", "cat3", cfg.bott + c3, u3, "dec3a", d3), ("dec3.3", "dec3a", d3, u3, "d3", d3),
            ("dec2.0", "cat2", d3 + c2, u2, "dec2a", d2), ("dec2.3", "dec2a", d2, u2, "d2", d2),
            ("dec1.0", "cat1", d2 + c1, u1, "dec1a", d1), ("dec1.3", "dec1a", d1, u1, "d1", d1),
        ]
        shapes = {
            "x0": (s0, c0p), "enc1a": (s0, c1), "e1": (s0, c1), "p1": (s1, c1),
            "enc2a": (s1, c2), "e2": (s1, c2), "p2": (s2, c2),
            "enc3a": (s2, c3), "e3": (s2, c3), "p3": (s3, c3), "b": (s3, cfg.bott),
            "cat3": (u3, cfg.bott + c3), "dec3a": (u3, d3), "d3": (u3, d3),
            "cat2": (u2, d3 + c2), "dec2a": (u2, d2), "d2": (u2, d2),
            "cat1": (u1, d2 + c1), "dec1a": (u1, d1), "d1": (u1, d1),
        }
        self.last = "d1"
        if cfg.final_resize:
            shapes["d1r"] = (s0, d1)
            self.last = "d1r"
        self.bufs = {k: buf(h, c) for k, (h, c) in shapes.items()}          # activations
        self.grads = {k: buf(h, c) for k, (h, c) in shapes.items()}         # dL/d(activation)
        # Pre-BN conv outputs y.  The three layers that consume a decoder concat [up | skip + time embedding] stay fp32: under the
        # per-sample embedding offsets (raw t <= 999) a bf16 y loses the spatial signal before the normalisation and flips ReLU
        # masks in the backward.  Everywhere else y is O(1..10) with no such offsets and is stored in the activation dtype
        # (bf16 on the tensor-core engine: what the reference's own conv output is under autocast); the batch statistics
        # still come from the fp32 accumulators in the conv epilogue.  TD_Y_FP32=1 keeps every y in fp32.
        all_fp32 = precision == "fp32" or os.environ.get("TD_Y_FP32", "0") != "0"
        self.yraw = {name: buf(size, cout, torch.float32 if (all_fp32 or xin.startswith("cat")) else self.act)
                     for name, xin, _, size, _, cout in self.layers}
        # dL/d(conv output), one buffer per layer: the weight gradients read them on a second stream while the main
        # stream has moved on to the next layers (151 MB at B = 128)
        self.dy = {name: buf(size, cout) for name, _, _, size, _, cout in self.layers}
        self.x_in = torch.zeros(B, cfg.in_ch, s0, s0, device=dev, dtype=torch.float32)
        self.eps = torch.zeros(B, cfg.in_ch, s0, s0, device=dev, dtype=torch.float32)
        self.d_eps = torch.zeros(B, cfg.in_ch, s0, s0, device=dev, dtype=torch.float32)
        self.temb = torch.zeros(B, c1 + c2 + c3, device=dev, dtype=torch.float32)
        self.d_temb = torch.zeros(B, c1 + c2 + c3, device=dev, dtype=torch.float32)
        self.emb_saved = torch.zeros(int(self.lib.td_embed_head_saved_floats(B, cfg.time_dim, cfg.emb_mode)),
                                     device=dev, dtype=torch.float32)
        self.emb_scratch = torch.zeros(2 * B * cfg.time_dim, device=dev, dtype=torch.float32)
        self.t_in = torch.zeros(B, device=dev, dtype=torch.int64)
        self.y_in = torch.zeros(B, device=dev, dtype=torch.int64) if cfg.cond == "class" else None
        self.text_in = (torch.zeros(B, cfg.time_dim, device=dev, dtype=torch.float32) if cfg.cond == "text" else None)

        # per-BN-layer state
        self.bn: Dict[str, Dict[str, torch.Tensor]] = {}
        max_part = 1
        for name, _, _, size, _, cout in self.layers:
            rows = max(int(self.lib.td_chan_reduce_rows(L.TD_F32, B * size * size, cout)),
                       int(self.lib.td_chan_reduce_rows(self.adt, B * size * size, cout)))
            rows_b = int(self.lib.td_bn_bwd_reduce_rows(self.adt, B * size * size, cout))
            rows_fused = B * size * size // 32 + 2          # upper bound on the conv epilogue's partial rows (tiles)
            max_part = max(max_part, (max(rows, rows_b, rows_fused) * 2 + 1) * cout)
            self.bn[name] = {k: torch.zeros(cout, device=dev) for k in ("scale", "shift", "mean", "invstd")}
            self.bn[name]["coef"] = torch.zeros(3, cout, device=dev)
            self.bn[name]["rows"] = rows
            self.bn[name]["rows_bwd"] = rows_b
        rows0 = int(self.lib.td_chan_reduce_rows(self.adt, B * s0 * s0, c0))
        max_part = max(max_part, (rows0 * 2 + 1) * c0, 128 * cfg.in_ch)
        self.rows_x0 = rows0
        self.partials = torch.zeros(max_part, device=dev)

        # parameter gradients (PyTorch layouts), keyed like state_dict
        self.pgrad: Dict[str, torch.Tensor] = {k: torch.zeros_like(p, device=dev) for k, p in module.named_parameters()}
        self.proj_w = torch.zeros(c1 + c2 + c3, cfg.time_dim, device=dev)
        self.proj_b = torch.zeros(c1 + c2 + c3, device=dev)
        self.d_proj_w = torch.zeros_like(self.proj_w)
        self.d_proj_b = torch.zeros_like(self.proj_b)

        self._alloc_packed()
        # forward convs (cout, size) and data-gradient convs (their "cout" is the layer's cin)
        self.splitk_ws = alloc_splitk_ws(self.lib, batch, [(cout, size) for _, _, _, size, _, cout in self.layers] +
                                         [(cin, size) for _, _, cin, size, _, _ in self.layers], device)
        self._build()
        self._weights_version = None
        self._pack_table = None

    # ------------------------------------------------------------------ parameters
    def _conv_modules(self):
        m = self.module
        out = [("initial_conv", m.initial_conv, None)]
        for blk, idx in _CONVS:
            seq = getattr(m, blk)
            out.append((f"{blk}.{idx}", seq[idx], seq[idx + 1]))
        out.append(("final_conv", m.final_conv, None))
        return out

    def _tc(self, cin: int, cout: int) -> bool:
        return self.precision == "bf16" and cin % 64 == 0 and cout % 64 == 0

    def _alloc_packed(self):
        self.w_fwd: Dict[str, torch.Tensor] = {}
        self.w_bwd: Dict[str, torch.Tensor] = {}
        self.conv_of: Dict[str, Tuple[torch.nn.Module, Optional[torch.nn.Module]]] = {}
        for name, conv, bn in self._conv_modules():
            cout, cin = conv.weight.shape[0], conv.weight.shape[1]
            if name == "enc1.0":
                cin = self.c0p
            self.conv_of[name] = (conv, bn)
            direct = name in ("initial_conv", "final_conv")
            wdt_f = torch.bfloat16 if (not direct and self._tc(cin, cout)) else torch.float32
            wdt_b = torch.bfloat16 if (not direct and self._tc(cout, cin)) else torch.float32
            self.w_fwd[name] = torch.zeros(cout, 3, 3, cin, device=self.device, dtype=wdt_f)
            if cin != conv.weight.shape[1]:
                self._w_pad_tmp = torch.zeros(cout, 3, 3, conv.weight.shape[1], device=self.device, dtype=wdt_f)
            if name != "initial_conv":
                self.w_bwd[name] = torch.zeros(cin, 3, 3, cout, device=self.device, dtype=wdt_b)

    def weights_version(self):
        # _weights_gen: bumped by TrainStep, whose kernels update the parameters through raw pointers (no _version bump)
        return (getattr(self.module, "_weights_gen", 0),) + tuple(p._version for p in self.module.parameters())

    def refresh_weights(self, force: bool = False) -> None:
        ver = self.weights_version()
        if not force and ver == self._weights_version:
            return
        st = L.stream_ptr()
        lib = self.lib
        if self._pack_table is None:
            self._build_pack_table()
        if self._pack_tiles > 0:          # every bf16 layer (forward + dgrad operand) in one launch
            L.check(lib.td_pack_conv_weights_multi(self._pack_table.data_ptr(), self._pack_entries, self._pack_tiles, st),
                    "td_pack_conv_weights_multi")
        for name, (conv, bn) in self.conv_of.items():
            if name in self._pack_fused:
                continue
            w = conv.weight.detach()
            assert w.is_cuda and w.dtype == torch.float32 and w.is_contiguous()
            pk = self.w_fwd[name]
            padded = pk.shape[3] != w.shape[1]           # enc1.0 with zero-padded input channels (c0p)
            if padded:
                pk = self._w_pad_tmp
            L.check(lib.td_pack_conv_weight(w.data_ptr(), pk.data_ptr(), L.dtype_code(pk.dtype), w.shape[0], w.shape[1],
                                            st), "td_pack_conv_weight")
            if padded:
                self.w_fwd[name][..., :w.shape[1]].copy_(pk)
            if name in self.w_bwd:
                # [cin][3][3][cout]: the real input channels are the leading rows of the padded operand
                pb = self.w_bwd[name]
                L.check(lib.td_pack_conv_weight_dgrad(w.data_ptr(), pb.data_ptr(), L.dtype_code(pb.dtype), w.shape[0],
                                                      w.shape[1], st), "td_pack_conv_weight_dgrad")
        m = self.module
        c1, c2, c3 = self.cfg.enc[1:]
        D = self.cfg.time_dim
        off = 0
        for proj, c in ((m.time_proj1, c1), (m.time_proj2, c2), (m.time_proj3, c3)):
            self.proj_w[off:off + c].copy_(proj.weight.detach().view(c, D))
            self.proj_b[off:off + c].copy_(proj.bias.detach())
            off += c
        self._weights_version = ver

    def _build_pack_table(self) -> None:
        """Device table for td_pack_conv_weights_multi: every conv whose packed operands are bf16 with channel counts
        that are multiples of 32 (the tensor-core layers); the parameter tensors are stable (updated in place)."""
        import struct
        rows, tiles, fused = [], 0, set()
        for name, (conv, bn) in self.conv_of.items():
            w = conv.weight
            pk = self.w_fwd[name]
            pb = self.w_bwd.get(name)
            co, ci = w.shape[0], w.shape[1]
            if pk.dtype != torch.bfloat16 or co % 32 or ci % 32 or (pb is not None and pb.dtype != torch.bfloat16):
                continue
            if pk.shape[3] != ci:
                continue
            assert w.is_cuda and w.dtype == torch.float32 and w.is_contiguous()
            rows.append(struct.pack("<QQQiiii", w.data_ptr(), pk.data_ptr(), pb.data_ptr() if pb is not None else 0,
                                    co, ci, tiles, 0))
            tiles += (co // 32) * (ci // 32)
            fused.add(name)
        self._pack_fused, self._pack_tiles, self._pack_entries = fused, tiles, len(rows)
        raw = b"".join(rows) if rows else b"\0" * 40
        self._pack_table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.device)

    # ------------------------------------------------------------------ plan construction
    def _conv_desc(self, x, cin, y, cout, w, x_coff=0, shift=None, x_nchw=False, y_nchw=False, size=None):
        d = L.ConvDesc()
        H = size if size is not None else (x.shape[2] if x_nchw else x.shape[1])
        d.batch, d.height, d.width, d.cin, d.cout = self.B, H, H, cin, cout
        d.x_dtype, d.y_dtype = L.dtype_code(x.dtype), L.dtype_code(y.dtype)
        d.x, d.ldx, d.x_coff = x.data_ptr(), (cin if x_nchw else x.shape[3]), x_coff
        d.y, d.ldy, d.y_coff = y.data_ptr(), (cout if y_nchw else y.shape[3]), 0
        d.w, d.scale, d.shift, d.relu, d.stats = w.data_ptr(), None, L.ptr(shift), 0, None
        d.x_nchw, d.y_nchw = int(x_nchw), int(y_nchw)
        d.splitk_ws = L.ptr(self.splitk_ws)
        return d

    def _wgrad(self, name, x, cin, dy, cout, size, engine, x_nchw=False, dy_nchw=False, dw=None):
        d = L.WgradDesc()
        d.batch, d.height, d.width, d.cin, d.cout = self.B, size, size, cin, cout
        d.x_dtype, d.dy_dtype = L.dtype_code(x.dtype), L.dtype_code(dy.dtype)
        d.x, d.ldx, d.x_coff, d.x_nchw = x.data_ptr(), (cin if x_nchw else x.shape[3]), 0, int(x_nchw)
        d.dy, d.lddy, d.dy_coff, d.dy_nchw = dy.data_ptr(), (cout if dy_nchw else dy.shape[3]), 0, int(dy_nchw)
        d.dw = (self.pgrad[self._wkey(name)] if dw is None else dw).data_ptr()
        need = int(self.lib.td_conv3x3_wgrad_workspace(C.byref(d), engine))
        self._wg_specs.append((name, d, engine, need))

    @staticmethod
    def _wkey(name: str) -> str:
        return f"{name}.weight"

    def _build(self):
        cfg, bf, gr, lib, B, S = self.cfg, self.bufs, self.grads, self.lib, self.B, self.S
        c0, c1, c2, c3 = cfg.enc
        d3, d2, d1 = cfg.dec
        adt = self.adt
        self.conv_plans: Dict[str, _ConvPlan] = {}
        self._wg_specs: List = []
        self._wg_post: Dict[str, Callable[[], None]] = {}
        self._side = getattr(self, "_side", None)
        self._side_f = getattr(self, "_side_f", None)
        fwd: List[Tuple[str, Callable[[int], None]]] = []
        bwd: List[Tuple[str, Callable[[int], None]]] = []
        m = self.module
        part = self.partials.data_ptr()

        def conv_plan(key, desc, engine):
            p = _ConvPlan(desc, engine)
            self.conv_plans[key] = p
            return p

        # ---- forward ---------------------------------------------------------------------------
        fwd.append(("embed", self._run_embed))
        ic = m.initial_conv
        p = conv_plan("initial_conv", self._conv_desc(self.x_in, cfg.in_ch, bf["x0"], c0, self.w_fwd["initial_conv"],
                                                      shift=ic.bias, x_nchw=True), L.CONV_DIRECT)
        fwd.append(("initial_conv", p.run))

        def add_block(name, xin, cin, size, out, cout):
            conv, bn = self.conv_of[name]
            x, y, a = bf[xin], self.yraw[name], bf[out]
            eng = L.CONV_TC if self._tc(cin, cout) else L.CONV_SIMT
            cd = self._conv_desc(x, cin, y, cout, self.w_fwd[name])
            if eng == L.CONV_TC:
                cd.stats = part                         # batch statistics from the fp32 accumulators (epilogue)
            p = conv_plan(name, cd, eng)
            fwd.append((name, p.run))
            st_ = self.bn[name]
            rows, P = int(lib.td_chan_reduce_rows(L.dtype_code(y.dtype), B * size * size, cout)), B * size * size
            fused_rows = int(lib.td_conv3x3_stats_rows(p.handle))
            assert (fused_rows * 2 + 1) * cout <= self.partials.numel()
            yp, ap = y.data_ptr(), a.data_ptr()
            g_, b_, cb = bn.weight.data_ptr(), bn.bias.data_ptr(), conv.bias.data_ptr()
            rm, rv, nbt = bn.running_mean.data_ptr(), bn.running_var.data_ptr(), bn.num_batches_tracked.data_ptr()
            eps_, mom = float(bn.eps), float(bn.momentum if bn.momentum is not None else 0.1)
            sc, sh, mu, iv = (st_[k].data_ptr() for k in ("scale", "shift", "mean", "invstd"))

            ydt = L.dtype_code(y.dtype)

            def bn_fwd(st):
                nrows = fused_rows
                if fused_rows == 0:
                    nrows = rows
                    L.check(lib.td_bn_stats(yp, ydt, cout, 0, P, cout, part, 1, st), "td_bn_stats")
                # finalize (partial rows -> scale / shift, running statistics) in the prologue of the apply + ReLU pass
                L.check(lib.td_bn_apply_fused(yp, ydt, part, nrows, P, g_, b_, cb, eps_, mom, rm, rv, nbt, sc, sh, mu, iv, ap, adt,
                                              cout, 0, P, cout, 1, st), "td_bn_apply_fused")
            fwd.append((f"bn:{name}", bn_fwd))

        def add_pool(src, dst, h, c):
            sp, dp = bf[src].data_ptr(), bf[dst].data_ptr()
            ceil = int(cfg.ceil_pool)
            fwd.append((f"pool:{src}", lambda st: L.check(lib.td_maxpool2_fwd(sp, dp, adt, B, h, h, c, ceil, st),
                                                          "td_maxpool2_fwd")))

        def add_upcat(low, skip, dst, ho, cu, hs, cs, toff):
            lp, sp, dp, tp = bf[low].data_ptr(), bf[skip].data_ptr(), bf[dst].data_ptr(), self.temb.data_ptr()
            ld = self.temb.shape[1]
            fwd.append((f"upcat:{dst}", lambda st: L.check(
                lib.td_upcat_fwd(lp, sp, tp, ld, toff, dp, adt, B, ho, ho, cu, hs, hs, cs, st), "td_upcat_fwd")))

        Ls = {l[0]: l for l in self.layers}
        for nm in ("enc1.0", "enc1.3"):
            add_block(*Ls[nm])
        add_pool("e1", "p1", S["s0"], c1)
        for nm in ("enc2.0", "enc2.3"):
            add_block(*Ls[nm])
        add_pool("e2", "p2", S["s1"], c2)
        for nm in ("enc3.0", "enc3.3"):
            add_block(*Ls[nm])
        add_pool("e3", "p3", S["s2"], c3)
        add_block(*Ls["bottleneck.0"])
        add_upcat("b", "e3", "cat3", S["u3"], cfg.bott, S["s2"], c3, c1 + c2)
        for nm in ("dec3.0", "dec3.3"):
            add_block(*Ls[nm])
        add_upcat("d3", "e2", "cat2", S["u2"], d3, S["s1"], c2, c1)
        for nm in ("dec2.0", "dec2.3"):
            add_block(*Ls[nm])
        add_upcat("d2", "e1", "cat1", S["u1"], d2, S["s0"], c1, 0)
        for nm in ("dec1.0", "dec1.3"):
            add_block(*Ls[nm])
        if cfg.final_resize:
            fwd.append(("resize:d1r", lambda st, sp=bf["d1"].data_ptr(), dp=bf["d1r"].data_ptr(), u1=S["u1"], s0=S["s0"]:
                        L.check(lib.td_resize_bilinear_fwd(sp, dp, adt, B, u1, u1, s0, s0, d1, st),
                                "td_resize_bilinear_fwd")))
        fc = m.final_conv
        p = conv_plan("final_conv", self._conv_desc(bf[self.last], d1, self.eps, cfg.in_ch, self.w_fwd["final_conv"],
                                                    shift=fc.bias, y_nchw=True), L.CONV_DIRECT)
        if cfg.in_ch == 1 and self.w_fwd["final_conv"].dtype == torch.float32:
            # Cout = 1: contract the channels once per pixel (nine scalars), then run the 3x3 stencil on scalars -- the
            # eval path's fused tail with an identity resize.  The direct kernel re-reads every pixel's channel vector for
            # each of the nine taps (35 us against ~15 us at B = 128).
            lp, wp, bp, ep = bf[self.last].data_ptr(), self.w_fwd["final_conv"].data_ptr(), fc.bias.data_ptr(), self.eps.data_ptr()
            hs = S["s0"] if cfg.final_resize else S["u1"]
            fwd.append(("final_conv", lambda st: L.check(
                lib.td_final_resize_conv(lp, adt, d1, 0, B, hs, hs, d1, wp, bp, hs, hs, ep, st), "td_final_resize_conv")))
        else:
            fwd.append(("final_conv", p.run))

        # ---- backward --------------------------------------------------------------------------
        s0 = S["s0"]
        # final_conv: bias / weight gradients and the data gradient (a tiny-Cin direct conv of d_eps)
        bwd.append(("final_conv:dbias", lambda st, dep=self.d_eps.data_ptr(), fcb=self.pgrad["final_conv.bias"].data_ptr():
                    L.check(lib.td_nchw_chansum(dep, B, cfg.in_ch, s0 * s0, fcb, part, st), "td_nchw_chansum")))
        self._wgrad("final_conv", bf[self.last], d1, self.d_eps, cfg.in_ch, s0, L.CONV_SIMT, dy_nchw=True)
        bwd.append(("final_conv:wgrad", None))
        p = conv_plan("final_conv:dgrad", self._conv_desc(self.d_eps, cfg.in_ch, gr[self.last], d1,
                                                          self.w_bwd["final_conv"], x_nchw=True), L.CONV_DIRECT)
        bwd.append(("final_conv:dgrad", p.run))
        if cfg.final_resize:
            bwd.append(("resize:d1r:bwd", lambda st, gp=gr["d1r"].data_ptr(), dp=gr["d1"].data_ptr(), u1=S["u1"]:
                        L.check(lib.td_resize_bilinear_bwd(gp, d1, 0, dp, adt, B, u1, u1, s0, s0, d1, st),
                                "td_resize_bilinear_bwd")))

        def add_block_bwd(name, xin, cin, size, out, cout, need_dx=True):
            conv, bn = self.conv_of[name]
            st_ = self.bn[name]
            rows, P = st_["rows_bwd"], B * size * size
            y, da = self.yraw[name], gr[out]
            dyv = self.dy[name]
            yp, dap, dyp = y.data_ptr(), da.data_ptr(), dyv.data_ptr()
            sc, sh, mu, iv = (st_[k].data_ptr() for k in ("scale", "shift", "mean", "invstd"))
            coef = st_["coef"].data_ptr()
            bt = bn.bias.data_ptr()
            ydt = L.dtype_code(y.dtype)
            blk, idx = name.split(".")
            dg = self.pgrad[f"{blk}.{int(idx) + 1}.weight"].data_ptr()
            db = self.pgrad[f"{blk}.{int(idx) + 1}.bias"].data_ptr()

            def bn_bwd(st):
                L.check(lib.td_bn_bwd_reduce(dap, cout, 0, yp, ydt, adt, sc, bt, mu, P, cout, part, st), "td_bn_bwd_reduce")
                # finalize (sum g, sum g*xhat -> dgamma / dbeta and the dy coefficients) in the prologue of the apply pass
                L.check(lib.td_bn_bwd_apply_fused(dap, cout, 0, yp, ydt, adt, part, rows, P, sc, bt, mu, iv, dg, db, dyp, P, cout, st),
                        "td_bn_bwd_apply_fused")
            bwd.append((f"bn:{name}:bwd", bn_bwd))
            eng = L.CONV_TC if self._tc(cin, cout) else L.CONV_SIMT
            wkey = self._wkey(name)
            if tuple(self.pgrad[wkey].shape) != (cout, cin, 3, 3):
                # zero-padded input channels (enc1.0 reading the widened x0): the kernel writes [cout][c0p][3][3] into a
                # scratch tensor and the real channels are copied into the flat gradient buffer on the same stream
                real = self.pgrad[wkey].shape[1]
                pad = torch.zeros(cout, cin, 3, 3, device=self.device, dtype=torch.float32)
                self._wg_post[name] = lambda pad=pad, real=real, wkey=wkey: self.pgrad[wkey].copy_(pad[:, :real])
                self._wgrad(name, bf[xin], cin, dyv, cout, size, eng, dw=pad)
            else:
                self._wgrad(name, bf[xin], cin, dyv, cout, size, eng)
            bwd.append((f"{name}:wgrad", None))
            if need_dx:
                engd = L.CONV_TC if self._tc(cout, cin) else L.CONV_SIMT
                p = conv_plan(f"{name}:dgrad", self._conv_desc(dyv, cout, gr[xin], cin, self.w_bwd[name], size=size), engd)
                bwd.append((f"{name}:dgrad", p.run))

        def add_upcat_bwd(low, skip, dst, ho, cu, hs, cs, toff):
            gp, lp, sp, tp = gr[dst].data_ptr(), gr[low].data_ptr(), gr[skip].data_ptr(), self.d_temb.data_ptr()
            ld = self.d_temb.shape[1]
            bwd.append((f"upcat:{dst}:bwd", lambda st: L.check(
                lib.td_upcat_bwd(gp, lp, sp, tp, ld, toff, adt, B, ho, ho, cu, hs, hs, cs, st), "td_upcat_bwd")))

        def add_pool_bwd(src, dst, h, c):
            xp, gyp, gxp = bf[src].data_ptr(), gr[dst].data_ptr(), gr[src].data_ptr()
            ceil = int(cfg.ceil_pool)
            bwd.append((f"pool:{src}:bwd", lambda st: L.check(
                lib.td_maxpool2_bwd(xp, gyp, gxp, adt, B, h, h, c, ceil, 1, st), "td_maxpool2_bwd")))

        for nm in ("dec1.3", "dec1.0"):
            add_block_bwd(*Ls[nm])
        add_upcat_bwd("d2", "e1", "cat1", S["u1"], d2, S["s0"], c1, 0)
        for nm in ("dec2.3", "dec2.0"):
            add_block_bwd(*Ls[nm])
        add_upcat_bwd("d3", "e2", "cat2", S["u2"], d3, S["s1"], c2, c1)
        for nm in ("dec3.3", "dec3.0"):
            add_block_bwd(*Ls[nm])
        add_upcat_bwd("b", "e3", "cat3", S["u3"], cfg.bott, S["s2"], c3, c1 + c2)
        # d_temb is complete here: the conditioning head's backward (a chain of small GEMMs) runs on a forked stream
        # beside the encoder backward and joins at its old place at the end of the plan (a parallel graph branch)
        cat3_name, cat3_fn = bwd[-1]

        def cat3_bwd_then_fork(st):
            cat3_fn(st)
            main = torch.cuda.current_stream()
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.device)
                self._ev_fork, self._ev_join = torch.cuda.Event(), torch.cuda.Event()
            self._ev_fork.record(main)
            self._side.wait_event(self._ev_fork)
            with torch.cuda.stream(self._side):
                self._run_embed_bwd(L.stream_ptr())
                self._ev_join.record(self._side)
        bwd[-1] = (cat3_name, cat3_bwd_then_fork)
        add_block_bwd(*Ls["bottleneck.0"])
        add_pool_bwd("e3", "p3", S["s2"], c3)
        for nm in ("enc3.3", "enc3.0"):
            add_block_bwd(*Ls[nm])
        add_pool_bwd("e2", "p2", S["s1"], c2)
        for nm in ("enc2.3", "enc2.0"):
            add_block_bwd(*Ls[nm])
        add_pool_bwd("e1", "p1", S["s0"], c1)
        for nm in ("enc1.3", "enc1.0"):
            add_block_bwd(*Ls[nm])
        # initial_conv: bias gradient = per-channel sum of d(x0); weight gradient; no data gradient
        gx0, icb, rows0 = gr["x0"].data_ptr(), self.pgrad["initial_conv.bias"].data_ptr(), self.rows_x0
        P0 = B * s0 * s0

        def ic_bias(st):
            L.check(lib.td_bn_stats(gx0, adt, self.c0p, 0, P0, c0, part, 0, st), "td_bn_stats")
            L.check(lib.td_partial_sum(part, rows0, c0, 0, icb, st), "td_partial_sum")
        bwd.append(("initial_conv:dbias", ic_bias))
        self._wgrad("initial_conv", self.x_in, cfg.in_ch, gr["x0"], c0, s0, L.CONV_SIMT, x_nchw=True)
        bwd.append(("initial_conv:wgrad", None))
        bwd.append(("embed:bwd", lambda st: torch.cuda.current_stream().wait_event(self._ev_join)))

        # Weight gradients run on side streams (parallel branches of the captured graph): nothing downstream in the
        # backward needs them, so the latency-bound BatchNorm / resize / pool backward kernels of the following layers
        # execute beside the tensor-core weight-gradient kernels instead of between them (2.63 -> 2.41 ms per step at
        # B = 128).  TD_WGRAD_STREAM=2 alternates consecutive layers between two side streams, each with its own split-K
        # workspace, so one layer's split-K reduction may overlap the next layer's main kernel; measured equal to one
        # stream (the main chain is the longer one), so one is the default; 0 keeps everything on the main stream.
        # Each entry forks from the main stream (its inputs are complete there) and the plan joins at the end.
        self.wgrad_streams = max(0, min(2, int(os.environ.get("TD_WGRAD_STREAM", "1"))))
        nws = max(1, self.wgrad_streams)
        ws_floats = max(need for _, _, _, need in self._wg_specs)
        self.wg_ws = [torch.zeros(max(ws_floats, 1), device=self.device) for _ in range(nws)]
        self.wg_plans: Dict[str, _WgradPlan] = {}
        self._wg_lane: Dict[str, int] = {}
        order = [n.split(":")[0] for n, fn in bwd if fn is None]          # execution order of the weight gradients
        for name, d, engine, _ in self._wg_specs:
            lane = order.index(name) % nws
            d.workspace = self.wg_ws[lane].data_ptr()
            self.wg_plans[name] = _WgradPlan(d, engine)
            self._wg_lane[name] = lane
        self._side_w = getattr(self, "_side_w", None) or [None] * 2
        self._wg_forked = [False, False]

        def on_wgrad_stream(name):
            run, lane = self.wg_plans[name].run, self._wg_lane[name]
            post = self._wg_post.get(name)
            if post is not None:
                plan_run = run

                def run(st):
                    plan_run(st)
                    post()
            if self.wgrad_streams == 0:
                return run

            def fn(st):
                main = torch.cuda.current_stream()
                if self._side_w[lane] is None:
                    self._side_w[lane] = torch.cuda.Stream(device=self.device)
                ev = torch.cuda.Event()
                ev.record(main)
                self._side_w[lane].wait_event(ev)
                with torch.cuda.stream(self._side_w[lane]):
                    run(L.stream_ptr())
                self._wg_forked[lane] = True
            return fn
        bwd = [(n, (on_wgrad_stream(n.split(":")[0]) if fn is None else fn)) for n, fn in bwd]
        bwd.append(("wgrad:join", lambda st: self.sync_wgrad_stream()))
        self.fwd_ops, self.bwd_ops = fwd, bwd

    def wgrad_side_stream(self):
        """The stream the weight gradients run on (None when they stay on the main stream or none was enqueued yet)."""
        if self.wgrad_streams != 1 or not self._wg_forked[0]:
            return None
        return self._side_w[0]

    def sync_wgrad_stream(self) -> None:
        """Make the current stream wait for every weight gradient enqueued so far (no-op when none is outstanding)."""
        for lane in range(2):
            if not self._wg_forked[lane]:
                continue
            ev = torch.cuda.Event()
            ev.record(self._side_w[lane])
            torch.cuda.current_stream().wait_event(ev)
            self._wg_forked[lane] = False

    # ------------------------------------------------------------------ conditioning head
    def _embed_args(self) -> "L.EmbedArgs":
        m, cfg = self.module, self.cfg
        a = L.EmbedArgs()
        a.batch, a.dim, a.in_mode, a.proj_out = self.B, cfg.time_dim, cfg.emb_mode, self.temb.shape[1]
        a.t, a.t_dev = self.t_in.data_ptr(), None
        mlp = m.time_mlp if cfg.cond == "text" else m.time_embedding
        a.w0, a.b0 = mlp[0].weight.data_ptr(), mlp[0].bias.data_ptr()
        a.w2, a.b2 = mlp[2].weight.data_ptr(), mlp[2].bias.data_ptr()
        a.y = self.y_in.data_ptr() if cfg.cond == "class" else None
        a.class_table = m.class_embedding.weight.data_ptr() if cfg.cond == "class" else None
        a.text = self.text_in.data_ptr() if cfg.cond == "text" else None
        a.proj_w, a.proj_b = self.proj_w.data_ptr(), self.proj_b.data_ptr()
        a.saved = self.emb_saved.data_ptr()
        a.proj_out_ptr = self.temb.data_ptr()
        return a

    def _run_embed(self, st: int) -> None:
        L.check(self.lib.td_embed_head_fwd(C.byref(self._embed_args()), st), "td_embed_head_fwd")

    def _run_embed_bwd(self, st: int) -> None:
        cfg = self.cfg
        pre = "time_mlp" if cfg.cond == "text" else "time_embedding"
        g = L.EmbedGrads()
        g.d_proj, g.scratch = self.d_temb.data_ptr(), self.emb_scratch.data_ptr()
        g.d_w0, g.d_b0 = self.pgrad[f"{pre}.0.weight"].data_ptr(), self.pgrad[f"{pre}.0.bias"].data_ptr()
        g.d_w2, g.d_b2 = self.pgrad[f"{pre}.2.weight"].data_ptr(), self.pgrad[f"{pre}.2.bias"].data_ptr()
        if cfg.cond == "class":
            g.d_class_table = self.pgrad["class_embedding.weight"].data_ptr()
            g.num_classes = self.module.class_embedding.weight.shape[0]
        g.d_proj_w, g.d_proj_b = self.d_proj_w.data_ptr(), self.d_proj_b.data_ptr()
        L.check(self.lib.td_embed_head_bwd(C.byref(self._embed_args()), C.byref(g), st), "td_embed_head_bwd")
        # scatter the fused projection gradient back to the three 1x1-conv parameters
        c1, c2, c3 = cfg.enc[1:]
        D, off = cfg.time_dim, 0
        for i, c in enumerate((c1, c2, c3), start=1):
            self.pgrad[f"time_proj{i}.weight"].view(c, D).copy_(self.d_proj_w[off:off + c])
            self.pgrad[f"time_proj{i}.bias"].copy_(self.d_proj_b[off:off + c])
            off += c

    # ------------------------------------------------------------------ execution
    def launch_forward(self) -> None:
        """The conditioning head (first entry) only feeds the decoder: it runs on a forked stream beside the encoder and
        joins before the first upcat (a parallel branch of the captured graph)."""
        st = L.stream_ptr()
        main = torch.cuda.current_stream()
        if self._side_f is None:
            self._side_f = torch.cuda.Stream(device=self.device)
            self._evf_fork, self._evf_join = torch.cuda.Event(), torch.cuda.Event()
        assert self.fwd_ops[0][0] == "embed"
        self._evf_fork.record(main)
        self._side_f.wait_event(self._evf_fork)
        with torch.cuda.stream(self._side_f):
            self.fwd_ops[0][1](L.stream_ptr())
            self._evf_join.record(self._side_f)
        joined = False
        for name, fn in self.fwd_ops[1:]:
            if not joined and name.startswith("upcat"):
                main.wait_event(self._evf_join)
                joined = True
            fn(st)
        if not joined:
            main.wait_event(self._evf_join)

    def launch_backward(self) -> None:
        """Reads d_eps; fills ``pgrad`` (every parameter).  Conv biases that feed a BatchNorm have a
        mathematically zero gradient in train mode and are left at zero."""
        st = L.stream_ptr()
        for _, fn in self.bwd_ops:
            fn(st)

    def load_inputs(self, x, t, cond) -> None:
        assert x.shape == self.x_in.shape, (x.shape, self.x_in.shape)
        self.x_in.copy_(x)
        self.t_in.copy_(t)
        if self.cfg.cond == "class":
            self.y_in.copy_(cond)
        elif self.cfg.cond == "text":
            self.text_in.copy_(cond)

    def num_launches(self) -> Tuple[int, int]:
        return len(self.fwd_ops), len(self.bwd_ops)

    def conv_flops(self) -> float:
        """Algorithmic FLOPs of one forward + backward (forward conv, data gradient, weight gradient)."""
        f = sum(p.flops for p in self.conv_plans.values())
        wg = sum(2.0 * d.batch * d.height * d.width * d.cout * 9.0 * d.cin for _, d, _, _ in self._wg_specs)
        return f + wg


class _UNetTrainFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine: UNetTrainEngine, x, t, cond, *params):
        engine.refresh_weights()
        engine.load_inputs(x, t, cond)
        engine.launch_forward()
        ctx.engine = engine
        ctx.n_params = len(params)
        # activations / BatchNorm statistics live in the engine's buffers, shared by every call at this batch size
        engine.generation = getattr(engine, "generation", 0) + 1
        ctx.generation = engine.generation
        return engine.eps.clone()

    @staticmethod
    def backward(ctx, d_eps):
        eng: UNetTrainEngine = ctx.engine
        _check_generation(eng, ctx)
        eng.d_eps.copy_(d_eps)
        eng.launch_backward()
        grads = tuple(eng.pgrad[k].clone() for k, _ in eng.module.named_parameters())
        return (None, None, None, None) + grads


def _check_generation(eng, ctx) -> None:
    if eng.generation != ctx.generation:
        raise RuntimeError(
            "tinydiff: backward() of a train-mode forward whose saved activations were overwritten by a later train-mode "
            f"forward at the same batch size (forward #{ctx.generation}, engine now at #{eng.generation}). The engine keeps "
            "one set of saved tensors per batch size: call backward() before the next forward (gradient accumulation over "
            "micro-batches: forward/backward one micro-batch at a time).")


def train_engine(model, batch: int, device: torch.device):
    if hasattr(model, "_declare"):                       # dense denoisers (latent MLP, DiT): dense.DenseEngine
        return model.engine(batch, device, training=True)
    key = ("train", batch, str(device), model.precision)
    eng = model._engines.get(key)
    if eng is None:
        cfg = model.config
        if cfg.time_dim != model.time_dim:
            cfg = UNetConfig(**{**cfg.__dict__, "time_dim": model.time_dim})
        eng = UNetTrainEngine(cfg, model, batch, device, model.precision)
        model._engines[key] = eng
    return eng


def unet_train_forward(model, x, t, cond):
    """Train-mode ``NoiseModel.forward`` with autograd (diffusion.py:228 followed by :235)."""
    device = x.device
    eng = train_engine(model, x.shape[0], device)
    params = tuple(p for _, p in model.named_parameters())
    return _UNetTrainFunction.apply(eng, x.to(torch.float32).contiguous(), t, cond, *params)


class _LRHandle(torch.optim.Optimizer):
    """A real ``torch.optim.Optimizer`` that owns nothing but the learning rate of a ``TrainStep``: the reference's
    scheduler lines (``CosineAnnealingLR(optimizer, T_max=num_epochs)`` stepped per epoch, diffusion_transformer.py:
    176-177,288; per batch with ``eta_min=1e-6``, conditional_diffusion_laion.py:434-438,473) work on it unchanged, and
    ``TrainStep`` reads ``param_groups[0]["lr"]`` before every step.  ``step()`` is a no-op (the fused kernel updates)."""

    def __init__(self, lr: float, device):
        self._dummy = torch.nn.Parameter(torch.zeros(1, device=device), requires_grad=False)
        super().__init__([self._dummy], {"lr": lr})

    def step(self, closure=None):       # noqa: D401
        return None


class TrainStep:
    """Fused train step: ``t ~ randint; x_t, noise = q_sample(x_0, t); eps = model(x_t, t[, y]);
    loss = mse(eps, noise); backward; [clip_grad_norm_;] Adam`` (diffusion.py:220-236;
    conditional_diffusion_laion.py:463-473) as one CUDA-graph replay.

    * ``lr`` lives in a device scalar read by the Adam kernel: ``set_lr()`` / the ``optimizer`` handle (a
      ``torch.optim.Optimizer`` for the reference's ``CosineAnnealingLR`` lines) change it under the captured graph.
    * ``max_grad_norm``: ``clip_grad_norm_(parameters, max_grad_norm)`` as one deterministic norm kernel whose clip
      factor feeds the Adam kernel's gradient scale (the clipped gradient is never materialised).
    * When neither ``t`` nor ``noise`` is injected both are drawn inside the graph (td_randint + td_qsample's Philox
      path): no library kernel runs in a step.
    * ``world_size > 1``: data parallel -- gradients are all-reduced (sum) through ``torch.distributed``
      (NCCL) in buckets (``dp.BucketReducer``), and the clip / Adam kernels scale them by 1/world_size.
    """

    def __init__(self, model, process, batch: int, device, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 use_graph: bool = True, process_group=None, bucket_mb: float = 8.0,
                 max_grad_norm: Optional[float] = None, seed: Optional[int] = None):
        self.device = L.require_device(device)
        self.model, self.process, self.B = model, process, batch
        self.lib = L.load()
        self.eng = train_engine(model, batch, self.device)
        self.betas, self.adam_eps = betas, eps
        self.use_graph = use_graph
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        self.max_grad_norm = None if max_grad_norm is None else float(max_grad_norm)
        dev = self.device
        e = self.eng
        self.x0 = torch.zeros_like(e.x_in)
        self.noise = torch.zeros_like(e.x_in)
        self.loss = torch.zeros(1, device=dev)
        n = e.eps.numel()
        self.partials = torch.zeros(int(self.lib.td_mse_num_partials(n)), device=dev)
        self.counter = torch.zeros(1, device=dev, dtype=torch.int32)
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int32)
        self.grad_scale = torch.full((1,), 1.0 / self.world, device=dev)
        self.grad_norm = torch.zeros(1, device=dev)              # total norm before clipping (max_grad_norm set)
        self.lr_dev = torch.full((1,), float(lr), device=dev)
        self._lr_host = float(lr)
        self.optimizer = _LRHandle(float(lr), dev)
        # device RNG of the step (t and the q_sample noise when nothing is injected): Philox key + per-step offset
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())   # one draw from the CPU generator (torch.manual_seed applies)
        self.rng = torch.tensor([seed, 0], dtype=torch.int64, device=dev)
        self.tab = process._tables(dev)
        # flat gradient buffer: pgrad tensors become views of it, so one all-reduce per bucket
        self.names = [k for k, _ in model.named_parameters()]
        self.params = [p for _, p in model.named_parameters()]
        sizes = [p.numel() for p in self.params]
        self.flat_grad = torch.zeros(sum((s + 3) // 4 * 4 for s in sizes), device=dev)
        off = 0
        self.offsets = []
        for k, p, s in zip(self.names, self.params, sizes):
            view = self.flat_grad[off:off + s].view_as(p)
            e.pgrad[k] = view
            self.offsets.append(off)
            off += (s + 3) // 4 * 4
        # data parallel: NCCL's all-reduce CTAs hold SMs while the backward runs; TD_DP_SM_RESERVE=n sizes the one-CTA-per-SM
        # grids for 148 - n SMs.  Off by default: with NCCL capped at 8 CTAs and n = 8 the 8-GPU step measured 2.30 ms against
        # 2.24 ms with NCCL's defaults and full grids (profiles/r02_dp_settings.txt).
        reserve = int(os.environ.get("TD_DP_SM_RESERVE", "0")) if self.world > 1 else 0
        prev_budget = self.lib.td_set_sm_budget(148 - reserve) if 0 < reserve <= 64 else None
        e._build()          # rebuild the plans against the flat gradient views
        if prev_budget is not None:
            self.lib.td_set_sm_budget(prev_budget)
        self.m = torch.zeros_like(self.flat_grad)
        self.v = torch.zeros_like(self.flat_grad)
        self.clip_partials = torch.zeros(int(self.lib.td_grad_clip_num_partials(self.flat_grad.numel())), device=dev)
        self.clip_counter = torch.zeros(1, device=dev, dtype=torch.int32)
        self._adam_tables()
        # gradient buckets: contiguous ranges of the flat buffer, each all-reduced as soon as the backward
        # entry that finalises its last gradient has been enqueued (overlaps the rest of the backward)
        ready = grad_ready_index(self.names, [n for n, _ in e.bwd_ops])
        self.buckets = plan_buckets(self.offsets, sizes, ready, int(bucket_mb * 1024 * 1024 / 4))
        self.reducer = BucketReducer(self.flat_grad, self.buckets, self.pg) if self.world > 1 else None
        self.graphs: Dict[bool, "torch.cuda.CUDAGraph"] = {}
        self.launches_per_step = 0
        self._device_rng = False
        self._ar_on_side = os.environ.get("TD_DP_AR_SIDE", "1") != "0"

    # -- learning rate -----------------------------------------------------------------------
    @property
    def lr(self) -> float:
        return self._lr_host

    @lr.setter
    def lr(self, value: float) -> None:
        self.set_lr(value)

    def set_lr(self, lr: float) -> None:
        """New learning rate from the next step on (device scalar: valid under the captured graph)."""
        lr = float(lr)
        self.optimizer.param_groups[0]["lr"] = lr
        self._push_lr(lr)

    def _push_lr(self, lr: float) -> None:
        if lr != self._lr_host:
            L.check(self.lib.td_fill_f32(self.lr_dev.data_ptr(), 1, lr, L.stream_ptr()), "td_fill_f32")
            self._lr_host = lr

    def _adam_tables(self):
        dev = self.device
        CH = 16384          # one CTA per 16 K elements: ~800 CTAs for the UNet, all resident at once (65536 left the pass latency-bound at 4.1 TB/s)
        ptrs_p, ptrs_g, ptrs_m, ptrs_v, numel, ct, co = [], [], [], [], [], [], []
        for i, (p, off) in enumerate(zip(self.params, self.offsets)):
            n = p.numel()
            ptrs_p.append(p.data_ptr())
            ptrs_g.append(self.flat_grad.data_ptr() + 4 * off)
            ptrs_m.append(self.m.data_ptr() + 4 * off)
            ptrs_v.append(self.v.data_ptr() + 4 * off)
            numel.append(n)
            for c in range(0, n, CH):
                ct.append(i)
                co.append(c)
        t64 = lambda v: torch.tensor(v, dtype=torch.int64, device=dev)
        self._tp, self._tg, self._tm, self._tv = t64(ptrs_p), t64(ptrs_g), t64(ptrs_m), t64(ptrs_v)
        self._numel = t64(numel)
        self._ct = torch.tensor(ct, dtype=torch.int32, device=dev)
        self._co = t64(co)
        self._chunks, self._chunk_elems = len(ct), CH
        self._ptrs = tuple(ptrs_p)

    def _check_pointers(self) -> None:
        """The plans and the Adam tables cache raw parameter pointers: refuse to run on storages that moved
        (``load_state_dict(assign=True)``, ``.to()``, ``p.data = ...``)."""
        if tuple(p.data_ptr() for p in self.params) != self._ptrs:
            raise RuntimeError("tinydiff.TrainStep: a parameter's storage changed since this TrainStep was built "
                               "(load_state_dict(assign=True) / .to() / .data assignment); build a new TrainStep")

    # -- pieces ------------------------------------------------------------------------------
    def _compute(self):
        """q_sample -> forward -> mse + dL/deps -> backward (everything up to the gradients)."""
        e, lib, st = self.eng, self.lib, L.stream_ptr()
        per = e.x_in.numel() // self.B
        seed_ptr = None
        if self._device_rng:                                       # diffusion.py:220 and :178 on the device, in the graph
            seed_ptr = self.rng.data_ptr()
            L.check(lib.td_randint(e.t_in.data_ptr(), self.B, 0, self.process.num_timesteps, seed_ptr, st), "td_randint")
        L.check(lib.td_qsample(self.x0.data_ptr(), self.noise.data_ptr(), e.t_in.data_ptr(), self.tab["abar"].data_ptr(),
                               e.x_in.data_ptr(), self.B, per, self.process.num_timesteps, seed_ptr, st), "td_qsample")
        if self._device_rng:
            L.check(lib.td_seed_advance(self.rng.data_ptr(), 1, st), "td_seed_advance")
        e.launch_forward()
        n = e.eps.numel()
        L.check(lib.td_mse_grad(e.eps.data_ptr(), self.noise.data_ptr(), e.d_eps.data_ptr(), self.loss.data_ptr(),
                                self.partials.data_ptr(), self.counter.data_ptr(), n, 1.0 / n, st), "td_mse_grad")
        if self.world == 1:
            e.launch_backward()
            return
        # data parallel: NCCL all-reduce of each bucket is enqueued (async, on NCCL's stream) right after the
        # backward entry that completes it, so the exchange overlaps the remaining backward kernels
        red = self.reducer
        for i, (_, fn) in enumerate(e.bwd_ops):
            fn(st)
            if not red.has_ready(i):
                continue
            # A bucket holds weight gradients (side stream) and BatchNorm / bias gradients (main stream).  Enqueue its
            # all-reduce from the side stream once that stream has also seen the main stream's progress: the main chain
            # never waits for a weight gradient here, only the final join before Adam does.
            ar_stream = e.wgrad_side_stream() if (self._ar_on_side and hasattr(e, "wgrad_side_stream")) else None
            if ar_stream is None:
                if hasattr(e, "sync_wgrad_stream"):
                    e.sync_wgrad_stream()
                red.enqueue_ready(i)
                continue
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            ar_stream.wait_event(ev)
            with torch.cuda.stream(ar_stream):
                red.enqueue_ready(i)

    def _allreduce(self):
        if self.reducer is not None:
            self.reducer.wait()

    def _update(self):
        lib, st = self.lib, L.stream_ptr()
        L.check(lib.td_counter_add(self.step_dev.data_ptr(), 1, st), "td_counter_add")
        if self.max_grad_norm is not None:                         # conditional_diffusion_laion.py:471
            L.check(lib.td_grad_clip_scale(self.flat_grad.data_ptr(), self.flat_grad.numel(), 1.0 / self.world,
                                           self.max_grad_norm, self.clip_partials.data_ptr(), self.clip_counter.data_ptr(),
                                           self.grad_scale.data_ptr(), self.grad_norm.data_ptr(), st), "td_grad_clip_scale")
        L.check(lib.td_adam_multi(self._tp.data_ptr(), self._tg.data_ptr(), self._tm.data_ptr(), self._tv.data_ptr(),
                                  self._numel.data_ptr(), self._ct.data_ptr(), self._co.data_ptr(), self._chunks,
                                  self._chunk_elems, self.step_dev.data_ptr(), self._lr_host, self.lr_dev.data_ptr(),
                                  self.betas[0], self.betas[1], self.adam_eps, self.grad_scale.data_ptr(), None, st),
                "td_adam_multi")
        self.eng.refresh_weights(force=True)

    def _body(self):
        self._compute()
        self._allreduce()
        self._update()

    # -- public ------------------------------------------------------------------------------
    def load(self, x_0, y=None, t=None, noise=None):
        """Stage one batch: x_0 (any device), optional labels / text, and optionally injected t / noise.  With neither
        injected, t and the noise are drawn on the device inside the step (``seed`` of the constructor)."""
        e = self.eng
        self.x0.copy_(x_0, non_blocking=True)
        if e.cfg.cond == "class":
            e.y_in.copy_(y, non_blocking=True)
        elif e.cfg.cond == "text":
            e.text_in.copy_(y, non_blocking=True)
        self._device_rng = t is None and noise is None
        if not self._device_rng:
            if t is None:
                e.t_in.copy_(torch.randint(0, self.process.num_timesteps, (self.B,), device=self.device))   # diffusion.py:220
            else:
                e.t_in.copy_(t, non_blocking=True)
            if noise is None:
                self.noise.normal_()                                                                    # diffusion.py:178
            else:
                self.noise.copy_(noise, non_blocking=True)
        if getattr(e, "_drop_slots", None):              # fresh dropout masks (outside the captured graph)
            e.reseed(int(torch.randint(0, 2 ** 62, (1,)).item()))

    def run(self) -> torch.Tensor:
        """One optimisation step on the staged batch; returns the (device) loss tensor."""
        self.model.train()
        self._check_pointers()
        self._push_lr(float(self.optimizer.param_groups[0]["lr"]))
        # the step writes parameters and BatchNorm running statistics through raw pointers: invalidate the packed
        # weights / folded BatchNorm of every eval-mode plan of this model (sample(), ValStep)
        self.model._weights_gen = getattr(self.model, "_weights_gen", 0) + 1
        if not self.use_graph:
            self.eng.refresh_weights()
            self._body()
            self.optimizer.step()
            return self.loss
        g = self.graphs.get(self._device_rng)
        if g is None:
            g = self._capture()
            self.graphs[self._device_rng] = g
        g.replay()
        self.optimizer.step()          # no-op: lets an attached lr_scheduler see "optimizer.step() before scheduler.step()"
        return self.loss

    def _capture(self) -> "torch.cuda.CUDAGraph":
        # world > 1: the bucketed NCCL all-reduces are captured too (every rank captures the same sequence; the
        # async works become cross-stream edges of the graph), so the data-parallel step is ONE replay per rank
        self.eng.refresh_weights(force=True)
        # warm-up outside capture on a side stream (lazy module load / cudaFuncSetAttribute), restoring
        # every piece of state the step mutates
        saved = [p.detach().clone() for p in self.params]
        bufs = [b.detach().clone() for b in self.model.buffers()]
        state = [t.clone() for t in (self.m, self.v, self.step_dev, self.rng, self.noise, self.eng.t_in)]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        n0 = int(self.lib.td_launch_count())
        with torch.cuda.graph(g):
            self._body()
        self.launches_per_step = int(self.lib.td_launch_count()) - n0         # library kernels in one train step
        with torch.no_grad():
            for p, s in zip(self.params, saved):
                p.copy_(s)
            for b, s in zip(self.model.buffers(), bufs):
                b.copy_(s)
            for t, s in zip((self.m, self.v, self.step_dev, self.rng, self.noise, self.eng.t_in), state):
                t.copy_(s)
        self.eng.refresh_weights(force=True)
        return g

    @property
    def graph(self):
        return next(iter(self.graphs.values()), None)

    def close(self) -> None:
        """Drop the captured graphs (with world_size > 1 they hold NCCL kernels: release them before the process group)."""
        self.graphs = {}

    def __call__(self, x_0, y=None, t=None, noise=None) -> torch.Tensor:
        self.load(x_0, y, t, noise)
        return self.run()


class ValStep:
    """Fused validation step (conditional_diffusion.py:275-289; conditional_diffusion_laion.py:497-519):
    ``t ~ randint; x_t, noise = q_sample(x_0, t); eps = model.eval()(x_t, t[, y]); loss = mse(eps, noise)`` with no
    gradient, as one CUDA-graph replay of the eval-mode plan (running-statistics BatchNorm folded into the conv
    epilogues, the kernels the sampler uses).  Like the reference's loop it leaves the model in eval mode; the
    caller switches back with ``model.train()`` (conditional_diffusion.py:351)."""

    def __init__(self, model, process, batch: int, device, use_graph: bool = True):
        self.device = L.require_device(device)
        self.model, self.process, self.B = model, process, batch
        self.lib = L.load()
        self.use_graph = use_graph
        self.eng = model.engine(batch, self.device)
        e = self.eng
        self.x0 = torch.zeros_like(e.x_in)
        self.noise = torch.zeros_like(e.x_in)
        self.loss = torch.zeros(1, device=self.device)
        n = e.eps.numel()
        self.partials = torch.zeros(int(self.lib.td_mse_num_partials(n)), device=self.device)
        self.counter = torch.zeros(1, device=self.device, dtype=torch.int32)
        self.tab = process._tables(self.device)
        self.graph = None
        self._version = None

    def _body(self):
        e, lib, st = self.eng, self.lib, L.stream_ptr()
        per = e.x_in.numel() // self.B
        L.check(lib.td_qsample(self.x0.data_ptr(), self.noise.data_ptr(), e.t_in.data_ptr(), self.tab["abar"].data_ptr(),
                               e.x_in.data_ptr(), self.B, per, self.process.num_timesteps, None, st), "td_qsample")
        e.launch()
        n = e.eps.numel()
        L.check(lib.td_mse_grad(e.eps.data_ptr(), self.noise.data_ptr(), None, self.loss.data_ptr(),
                                self.partials.data_ptr(), self.counter.data_ptr(), n, 1.0 / n, st), "td_mse_grad")

    def load(self, x_0, y=None, t=None, noise=None):
        e = self.eng
        self.x0.copy_(x_0, non_blocking=True)
        if e.cfg.cond == "class":
            e.y_in.copy_(y, non_blocking=True)
        elif e.cfg.cond == "text":
            e.text_in.copy_(y, non_blocking=True)
        if t is None:
            e.t_in.copy_(torch.randint(0, self.process.num_timesteps, (self.B,), device=self.device))
        else:
            e.t_in.copy_(t, non_blocking=True)
        if noise is None:
            self.noise.normal_()
        else:
            self.noise.copy_(noise, non_blocking=True)

    @torch.no_grad()
    def run(self) -> torch.Tensor:
        self.model.eval()
        e = self.eng
        e.use_t_dev = False
        e.refresh_weights()          # re-folds BatchNorm / re-packs only when a parameter or buffer changed
        if not self.use_graph:
            self._body()
            return self.loss
        if self.graph is None:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._body()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._body()
            self.graph = g
        self.graph.replay()
        return self.loss

    def __call__(self, x_0, y=None, t=None, noise=None) -> torch.Tensor:
        self.load(x_0, y, t, noise)
        return self.run()
