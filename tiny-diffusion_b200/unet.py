"""Host-side plan of the conv-UNet denoiser (the reference's ``NoiseModel.forward``,
diffusion.py:109-162 / conditional_diffusion.py:112-171 / conditional_diffusion_laion.py:303-332)
expressed as a fixed sequence of libtinydiff kernel launches over preallocated NHWC buffers.

PyTorch owns the memory and the stream; every arithmetic step is a C-ABI call.  A plan is built
for one (batch, precision) pair and can be replayed inside a CUDA graph.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Tuple

import torch

from . import _lib as L


@dataclass(frozen=True)
class UNetConfig:
    """Static shape description of one reference conv UNet."""
    name: str
    in_ch: int
    enc: Tuple[int, int, int, int]       # initial_conv, enc1, enc2, enc3 output channels
    bott: int
    dec: Tuple[int, int, int]            # dec3, dec2, dec1 output channels
    time_dim: int
    ceil_pool: bool                      # MaxPool2d(2, ceil_mode=True)   diffusion.py:101
    resize_skips: bool                   # F.interpolate(e_k + t_k, size=...)  diffusion.py:137-153
    final_resize: bool                   # interpolate d1 back to the input size  diffusion.py:157-159
    emb_mode: int                        # 0 raw t, 2 sinusoidal (td_embed_args.in_mode)
    cond: str                            # "none" | "class" | "text"
    image: int


MNIST_UNET = UNetConfig("diffusion", 1, (64, 128, 256, 512), 512, (256, 128, 64), 256, True, True, True, 0,
                        "none", 28)
COND_UNET = UNetConfig("conditional_diffusion", 1, (64, 128, 256, 512), 512, (256, 128, 64), 256, True, True,
                       True, 0, "class", 28)
LAION_UNET = UNetConfig("conditional_diffusion_laion", 4, (32, 64, 128, 256), 256, (256, 128, 64), 768, False,
                        False, False, 2, "text", 32)

# (block name, index of conv inside the nn.Sequential, index of its BatchNorm)
_CONVS = [("enc1", 0), ("enc1", 3), ("enc2", 0), ("enc2", 3), ("enc3", 0), ("enc3", 3), ("bottleneck", 0),
          ("dec3", 0), ("dec3", 3), ("dec2", 0), ("dec2", 3), ("dec1", 0), ("dec1", 3)]


def _pool(n: int, ceil: bool) -> int:
    return (n + 1) // 2 if ceil else n // 2


def alloc_splitk_ws(lib, batch: int, layers, device) -> Optional[torch.Tensor]:
    """One split-K workspace shared by every tcgen05 conv of a plan (layers: (cout, size) pairs)."""
    need = 0
    for cout, size in layers:
        d = L.ConvDesc()
        d.batch, d.height, d.width, d.cin, d.cout = batch, size, size, 64, cout
        need = max(need, int(lib.td_conv3x3_splitk_workspace(C.byref(d))))
    return torch.zeros(need, device=device) if need > 0 else None


class _ConvPlan:
    """Owns one td_conv_plan handle."""

    def __init__(self, desc: L.ConvDesc, engine: int):
        self.lib = L.load()
        self.handle = C.c_void_p()
        self.desc = desc
        self.engine = engine
        L.check(self.lib.td_conv3x3_plan_create(C.byref(self.handle), C.byref(desc), engine), "td_conv3x3_plan_create")
        self.flops = float(self.lib.td_conv3x3_flops(self.handle))

    def run(self, stream: int) -> None:
        L.check(self.lib.td_conv3x3_run(self.handle, stream), "td_conv3x3_run")

    def __del__(self):
        try:
            if self.handle:
                self.lib.td_conv3x3_plan_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass


class UNetEngine:
    """Eval-mode (running-statistics BatchNorm) forward plan for a fixed batch size.

    precision "bf16": NHWC bf16 activations, tcgen05 implicit-GEMM convolutions with fp32
    accumulation, fp32 conditioning head and fp32 network input/output.
    precision "fp32": NHWC fp32 activations, fp32 FFMA convolutions (parity path, 1e-4 budget).
    """

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
        if not cfg.resize_skips:
            assert (u3, u2, u1) == (s2, s1, s0), "skip sizes must match without resize"
        self.sizes = dict(s0=s0, s1=s1, s2=s2, s3=s3, u3=u3, u2=u2, u1=u1)
        B = batch
        # enc1.0 reads initial_conv's output.  When that has fewer than 64 channels (the LAION latent UNet: 32) the layer
        # would miss the tcgen05 engine (K tiles of 64 channels), so the buffer is allocated 64 wide, the upper channels
        # stay zero (initial_conv writes the first c0 with row stride c0p) and the packed weight is zero-padded to match.
        self.c0p = c0
        if precision == "bf16" and c0 % 64 != 0 and c1 % 64 == 0 and c0 % 8 == 0:
            self.c0p = (c0 + 63) // 64 * 64

        def buf(h, c, dtype=None):
            return torch.zeros(B, h, h, c, device=device, dtype=dtype or self.act)

        self.bufs: Dict[str, torch.Tensor] = {
            "x0": buf(s0, self.c0p),
            "enc1a": buf(s0, c1), "e1": buf(s0, c1), "p1": buf(s1, c1),
            "enc2a": buf(s1, c2), "e2": buf(s1, c2), "p2": buf(s2, c2),
            "enc3a": buf(s2, c3), "e3": buf(s2, c3), "p3": buf(s3, c3),
            "b": buf(s3, cfg.bott),
            "cat3": buf(u3, cfg.bott + c3), "dec3a": buf(u3, d3), "d3": buf(u3, d3),
            "cat2": buf(u2, d3 + c2), "dec2a": buf(u2, d2), "d2": buf(u2, d2),
            "cat1": buf(u1, d2 + c1), "dec1a": buf(u1, d1), "d1": buf(u1, d1),
        }
        if cfg.final_resize and cfg.in_ch != 1:
            self.bufs["d1r"] = buf(s0, d1)
        self.x_in = torch.zeros(B, cfg.in_ch, s0, s0, device=device, dtype=torch.float32)
        self.eps = torch.zeros(B, cfg.in_ch, s0, s0, device=device, dtype=torch.float32)
        self.temb = torch.zeros(B, c1 + c2 + c3, device=device, dtype=torch.float32)
        self.emb_saved = torch.zeros(int(self.lib.td_embed_head_saved_floats(B, cfg.time_dim, cfg.emb_mode)),
                                     device=device, dtype=torch.float32)
        self.t_in = torch.zeros(B, device=device, dtype=torch.int64)
        self.t_dev = torch.zeros(1, device=device, dtype=torch.int32)       # sampler step counter
        self.y_in = torch.zeros(B, device=device, dtype=torch.int64) if cfg.cond == "class" else None
        self.text_in = (torch.zeros(B, cfg.time_dim, device=device, dtype=torch.float32)
                        if cfg.cond == "text" else None)

        # packed parameters (refreshed by refresh_weights)
        self.packed: Dict[str, torch.Tensor] = {}
        self.scale: Dict[str, torch.Tensor] = {}
        self.shift: Dict[str, torch.Tensor] = {}
        self.proj_w = torch.zeros(c1 + c2 + c3, cfg.time_dim, device=device, dtype=torch.float32)
        self.proj_b = torch.zeros(c1 + c2 + c3, device=device, dtype=torch.float32)
        self._alloc_packed()
        self.splitk_ws = alloc_splitk_ws(self.lib, batch, [(c, s) for c, s in (
            (c1, s0), (c2, s1), (c3, s2), (cfg.bott, s3), (d3, u3), (d2, u2), (d1, u1))], device)
        self._build_plans()
        self._weights_version = None
        self._side = None

    # ------------------------------------------------------------------ parameters
    def _conv_modules(self):
        m = self.module
        out = [("initial_conv", m.initial_conv, None)]
        for blk, idx in _CONVS:
            seq = getattr(m, blk)
            out.append((f"{blk}.{idx}", seq[idx], seq[idx + 1]))
        out.append(("final_conv", m.final_conv, None))
        return out

    def _engine_for(self, name: str, cin: int, cout: int) -> int:
        if name in ("initial_conv", "final_conv"):
            return L.CONV_DIRECT
        if self.precision == "bf16" and cin % 64 == 0 and cout % 64 == 0:
            return L.CONV_TC
        return L.CONV_SIMT

    def _alloc_packed(self):
        self.engines: Dict[str, int] = {}
        self._pack_tmp: Dict[str, torch.Tensor] = {}
        for name, conv, bn in self._conv_modules():
            cout, cin = conv.weight.shape[0], conv.weight.shape[1]
            cin_p = self.c0p if name == "enc1.0" else cin
            eng = self._engine_for(name, cin_p, cout)
            self.engines[name] = eng
            wdt = torch.bfloat16 if eng == L.CONV_TC else torch.float32
            self.packed[name] = torch.zeros(cout, 3, 3, cin_p, device=self.device, dtype=wdt)
            if cin_p != cin:
                self._pack_tmp[name] = torch.zeros(cout, 3, 3, cin, device=self.device, dtype=wdt)
            self.scale[name] = torch.ones(cout, device=self.device, dtype=torch.float32)
            self.shift[name] = torch.zeros(cout, device=self.device, dtype=torch.float32)

    def weights_version(self):
        # _weights_gen: bumped by train.TrainStep, whose kernels update parameters and BatchNorm running statistics through
        # raw pointers (PyTorch's _version counters do not see those writes)
        return (getattr(self.module, "_weights_gen", 0),) + tuple(p._version for p in self.module.parameters()) + tuple(
            b._version for b in self.module.buffers())

    def refresh_weights(self, force: bool = False) -> None:
        """Re-pack conv weights (OIHW fp32 -> OHWI bf16/fp32), fold eval-mode BatchNorm into the
        per-channel epilogue affine, and concatenate the three time_proj 1x1 convs."""
        ver = self.weights_version()
        if not force and ver == self._weights_version:
            return
        st = L.stream_ptr()
        lib = self.lib
        for name, conv, bn in self._conv_modules():
            w = conv.weight.detach()
            assert w.is_cuda and w.dtype == torch.float32 and w.is_contiguous()
            pk = self._pack_tmp.get(name, self.packed[name])
            L.check(lib.td_pack_conv_weight(w.data_ptr(), pk.data_ptr(), L.dtype_code(pk.dtype), w.shape[0],
                                            w.shape[1], st), "td_pack_conv_weight")
            if pk is not self.packed[name]:
                self.packed[name][..., :w.shape[1]].copy_(pk)      # zero-padded input channels (see c0p)
            if bn is None:
                self.shift[name].copy_(conv.bias.detach())
            else:
                L.check(lib.td_bn_fold(bn.weight.data_ptr(), bn.bias.data_ptr(), bn.running_mean.data_ptr(),
                                       bn.running_var.data_ptr(), conv.bias.data_ptr(), float(bn.eps),
                                       self.scale[name].data_ptr(), self.shift[name].data_ptr(), w.shape[0], st),
                        "td_bn_fold")
        m = self.module
        c1, c2, c3 = self.cfg.enc[1:]
        D = self.cfg.time_dim
        off = 0
        for proj, c in ((m.time_proj1, c1), (m.time_proj2, c2), (m.time_proj3, c3)):
            self.proj_w[off:off + c].copy_(proj.weight.detach().view(c, D))
            self.proj_b[off:off + c].copy_(proj.bias.detach())
            off += c
        self._weights_version = ver

    # ------------------------------------------------------------------ plan
    def _conv(self, name, x, cin, y, cout, relu, x_coff=0, y_coff=0, x_nchw=False, y_nchw=False, y_dtype=None,
              x_dtype=None, pool=None):
        B, H = self.B, (x.shape[2] if x_nchw else x.shape[1])
        d = L.ConvDesc()
        d.batch, d.height, d.width, d.cin, d.cout = B, H, H, cin, cout
        d.x_dtype = L.dtype_code(x.dtype) if x_dtype is None else x_dtype
        d.y_dtype = L.dtype_code(y.dtype) if y_dtype is None else y_dtype
        d.x, d.ldx, d.x_coff = x.data_ptr(), (cin if x_nchw else x.shape[3]), x_coff
        d.y, d.ldy, d.y_coff = y.data_ptr(), (cout if y_nchw else y.shape[3]), y_coff
        d.w = self.packed[name].data_ptr()
        has_bn = name not in ("initial_conv", "final_conv")
        d.scale = self.scale[name].data_ptr() if has_bn else None
        d.shift = self.shift[name].data_ptr()
        d.relu = 1 if relu else 0
        d.stats = None
        d.x_nchw, d.y_nchw = int(x_nchw), int(y_nchw)
        d.splitk_ws = L.ptr(self.splitk_ws)
        if pool is not None:                  # MaxPool2d(2) of y written by the conv epilogue when the plan accepts it
            d.pool_y, d.pool_ceil = pool.data_ptr(), int(self.cfg.ceil_pool)
        plan = _ConvPlan(d, self.engines[name])
        self.plans[name] = plan
        return plan

    def _build_plans(self):
        cfg, bf, lib, B = self.cfg, self.bufs, self.lib, self.B
        c0, c1, c2, c3 = cfg.enc
        d3, d2, d1 = cfg.dec
        S = self.sizes
        self.plans: Dict[str, _ConvPlan] = {}
        ops: List[Tuple[str, Callable[[int], None]]] = []

        def add_conv(name, *a, **k):
            plan = self._conv(name, *a, **k)
            ops.append((name, plan.run))
            return plan

        def add_conv_pool(name, x, cin, y_name, cout, p_name, h):
            """conv + BatchNorm + ReLU followed by MaxPool2d(2): one launch when the tcgen05 halo epilogue can pool (windows
            inside a subtile), else the separate pooling kernel."""
            fuse = self.adt == L.TD_BF16 and self.engines[name] == L.CONV_TC
            plan = add_conv(name, x, cin, bf[y_name], cout, True, pool=bf[p_name] if fuse else None)
            if not (fuse and int(lib.td_conv3x3_pool_fused(plan.handle))):
                add_pool(y_name, p_name, h, cout)

        def add_pool(src, dst, h, c):
            sp, dp = bf[src].data_ptr(), bf[dst].data_ptr()
            ceil = int(cfg.ceil_pool)
            adt = self.adt
            ops.append((f"pool:{src}", lambda st: L.check(lib.td_maxpool2_fwd(sp, dp, adt, B, h, h, c, ceil, st),
                                                          "td_maxpool2_fwd")))

        def add_upcat(low, skip, dst, ho, cu, hs, cs, toff):
            lp, sp, dp, tp = bf[low].data_ptr(), bf[skip].data_ptr(), bf[dst].data_ptr(), self.temb.data_ptr()
            ld = self.temb.shape[1]
            adt = self.adt
            ops.append((f"upcat:{dst}", lambda st: L.check(
                lib.td_upcat_fwd(lp, sp, tp, ld, toff, dp, adt, B, ho, ho, cu, hs, hs, cs, st), "td_upcat_fwd")))

        ops.append(("embed", self._run_embed))
        add_conv("initial_conv", self.x_in, cfg.in_ch, bf["x0"], c0, False, x_nchw=True)
        add_conv("enc1.0", bf["x0"], self.c0p, bf["enc1a"], c1, True)
        add_conv_pool("enc1.3", bf["enc1a"], c1, "e1", c1, "p1", S["s0"])
        add_conv("enc2.0", bf["p1"], c1, bf["enc2a"], c2, True)
        add_conv_pool("enc2.3", bf["enc2a"], c2, "e2", c2, "p2", S["s1"])
        add_conv("enc3.0", bf["p2"], c2, bf["enc3a"], c3, True)
        add_conv_pool("enc3.3", bf["enc3a"], c3, "e3", c3, "p3", S["s2"])
        add_conv("bottleneck.0", bf["p3"], c3, bf["b"], cfg.bott, True)
        add_upcat("b", "e3", "cat3", S["u3"], cfg.bott, S["s2"], c3, c1 + c2)
        add_conv("dec3.0", bf["cat3"], cfg.bott + c3, bf["dec3a"], d3, True)
        add_conv("dec3.3", bf["dec3a"], d3, bf["d3"], d3, True)
        add_upcat("d3", "e2", "cat2", S["u2"], d3, S["s1"], c2, c1)
        add_conv("dec2.0", bf["cat2"], d3 + c2, bf["dec2a"], d2, True)
        add_conv("dec2.3", bf["dec2a"], d2, bf["d2"], d2, True)
        add_upcat("d2", "e1", "cat1", S["u1"], d2, S["s0"], c1, 0)
        add_conv("dec1.0", bf["cat1"], d2 + c1, bf["dec1a"], d1, True)
        add_conv("dec1.3", bf["dec1a"], d1, bf["d1"], d1, True)
        if cfg.final_resize and cfg.in_ch == 1:
            # resize + final_conv in one pass over d1 (both are linear: the 64-channel resized tensor never exists)
            sp, ep = bf["d1"].data_ptr(), self.eps.data_ptr()
            wp, bp = self.packed["final_conv"].data_ptr(), self.shift["final_conv"].data_ptr()
            u1, s0, adt = S["u1"], S["s0"], self.adt
            ops.append(("final_resize_conv", lambda st: L.check(
                lib.td_final_resize_conv(sp, adt, d1, 0, B, u1, u1, d1, wp, bp, s0, s0, ep, st), "td_final_resize_conv")))
        else:
            last = "d1"
            if cfg.final_resize:
                sp, dp = bf["d1"].data_ptr(), bf["d1r"].data_ptr()
                u1, s0, adt = S["u1"], S["s0"], self.adt
                ops.append(("resize:d1r", lambda st: L.check(
                    lib.td_resize_bilinear_fwd(sp, dp, adt, B, u1, u1, s0, s0, d1, st), "td_resize_bilinear_fwd")))
                last = "d1r"
            add_conv("final_conv", bf[last], d1, self.eps, cfg.in_ch, False, y_nchw=True)
        self.ops = ops
        self.use_t_dev = False

    def _embed_args(self) -> "L.EmbedArgs":
        m, cfg = self.module, self.cfg
        a = L.EmbedArgs()
        a.batch, a.dim, a.in_mode, a.proj_out = self.B, cfg.time_dim, cfg.emb_mode, self.temb.shape[1]
        a.t = None if self.use_t_dev else self.t_in.data_ptr()
        a.t_dev = self.t_dev.data_ptr()
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
        if self.use_t_dev and getattr(self, "_pc_ready", False):
            # sampler: every sample shares t, so the time MLP runs once (batch 1, GEMV kernels) and the per-sample
            # conditioning enters through the linearity of the projection:
            #   proj_b = P (e_t + c_b) + pb = (P e_t + pb) + P c_b,   P c_b precomputed by prepare_sampler_embed()
            L.check(self.lib.td_embed_head_fwd(C.byref(self._embed_args1), st), "td_embed_head_fwd")
            L.check(self.lib.td_gemm_f32(C.byref(self._bcast_args), st), "td_gemm_f32")
            return
        L.check(self.lib.td_embed_head_fwd(C.byref(self._embed_args()), st), "td_embed_head_fwd")

    def _gemm_args(self, M, N, K, A, a_rs, a_cs, Bm, b_rs, b_cs, Cm, ldc):
        g = L.GemmArgs()
        g.M, g.N, g.K, g.alpha = M, N, K, 1.0
        g.A, g.a_rs, g.a_cs = A, a_rs, a_cs
        g.B, g.b_rs, g.b_cs = Bm, b_rs, b_cs
        g.C, g.ldc = Cm, ldc
        return g

    def prepare_sampler_embed(self) -> None:
        """Once per sample() call (labels / text embeddings and weights are fixed over the T steps)."""
        cfg, m, B = self.cfg, self.module, self.B
        P, D = self.temb.shape[1], cfg.time_dim
        st = L.stream_ptr()
        if not hasattr(self, "_v1"):
            dev = self.device
            self._v1 = torch.zeros(1, P, device=dev)
            self._ones = torch.ones(B, 1, device=dev)
            self._saved1 = torch.zeros(int(self.lib.td_embed_head_saved_floats(1, D, cfg.emb_mode)), device=dev)
            self._pc = torch.zeros(max(B, m.class_embedding.weight.shape[0]) if cfg.cond == "class" else B, P, device=dev)
        a = self._embed_args()
        a.batch, a.t, a.y, a.class_table, a.text = 1, None, None, None, None
        a.saved, a.proj_out_ptr = self._saved1.data_ptr(), self._v1.data_ptr()
        self._embed_args1 = a
        # temb[b, :] = 1 * v[:] (+ Pc[y_b, :] | + Pcond[b, :])
        g = self._gemm_args(B, P, 1, self._ones.data_ptr(), 1, 1, self._v1.data_ptr(), P, 1, self.temb.data_ptr(), P)
        if cfg.cond == "class":
            tab = m.class_embedding.weight
            pc = self._gemm_args(tab.shape[0], P, D, tab.data_ptr(), D, 1, self.proj_w.data_ptr(), 1, D,
                                 self._pc.data_ptr(), P)
            L.check(self.lib.td_gemm_f32(C.byref(pc), st), "td_gemm_f32")
            g.gather_idx, g.gather_table, g.ld_table = self.y_in.data_ptr(), self._pc.data_ptr(), P
        elif cfg.cond == "text":
            pc = self._gemm_args(B, P, D, self.text_in.data_ptr(), D, 1, self.proj_w.data_ptr(), 1, D,
                                 self._pc.data_ptr(), P)
            L.check(self.lib.td_gemm_f32(C.byref(pc), st), "td_gemm_f32")
            g.residual, g.ldr = self._pc.data_ptr(), P
        self._bcast_args = g
        self._pc_ready = True

    # ------------------------------------------------------------------ execution
    def launch(self) -> None:
        """Enqueue one eval forward on the current stream (graph-capturable): reads x_in / t_in (or
        t_dev) / y_in / text_in, writes eps.  The conditioning head only feeds the decoder (first use:
        the first upcat), so it runs on a forked stream beside the encoder convolutions and joins
        there -- in a captured graph that is a parallel branch, off the critical path."""
        main = torch.cuda.current_stream()
        st = L.stream_ptr()
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
            self._ev_fork, self._ev_join = torch.cuda.Event(), torch.cuda.Event()
        self._ev_fork.record(main)
        self._side.wait_event(self._ev_fork)
        with torch.cuda.stream(self._side):
            self.ops[0][1](L.stream_ptr())
            self._ev_join.record(self._side)
        joined = False
        for name, fn in self.ops[1:]:
            if not joined and name.startswith("upcat"):
                main.wait_event(self._ev_join)
                joined = True
            fn(st)
        if not joined:
            main.wait_event(self._ev_join)

    def num_launches(self) -> int:
        return len(self.ops)

    def conv_flops(self) -> float:
        return sum(p.flops for p in self.plans.values())

    def forward(self, x: torch.Tensor, t: torch.Tensor, cond: Optional[torch.Tensor] = None) -> torch.Tensor:
        assert x.shape == self.x_in.shape, (x.shape, self.x_in.shape)
        self.refresh_weights()
        self.x_in.copy_(x)
        self.t_in.copy_(t)
        if self.cfg.cond == "class":
            self.y_in.copy_(cond)
        elif self.cfg.cond == "text":
            self.text_in.copy_(cond)
        self.use_t_dev = False
        self.launch()
        return self.eps.clone()
