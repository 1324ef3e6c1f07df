"""Dense (fully connected) denoisers on libtinydiff: the latent MLP "U-Net"
(latent_diffusion.py:16-128) and the length-1-sequence "DiT" (diffusion_transformer.py:16-109).

``DenseEngine`` is a small static tape: the model's forward is declared once as a list of
libtinydiff launches over preallocated fp32 [batch, features] buffers (column slices of a buffer
stand in for ``torch.cat``), and the backward plan is derived from it in reverse order -- every
gradient lands either by overwrite (first producer in execution order) or through the GEMM
epilogue's accumulate flag (later producers).  It exposes the same surface as
``train.UNetTrainEngine`` so ``TrainStep``, the autograd wrapper and ``ReverseLoop`` work unchanged.

Arithmetic: fp32 FFMA GEMMs (`td_gemm_f32`), LayerNorm / BatchNorm1d / activation / dropout kernels.
These models are 2.8 / 5.4 MFLOP per sample -- latency-bound at the reference batch sizes.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, Dict, List, Optional, Tuple

import torch

from . import _lib as L
from .checkpoint import CheckpointCompat

Mat = Tuple[str, int, int]            # (buffer name, first column, end column)
_P = C.c_void_p


class TapeOp(C.Structure):
    """One op of the fused eval-mode forward (csrc/dense_fused.cu ``TapeOp``)."""
    _fields_ = [("kind", C.c_int), ("N", C.c_int), ("K", C.c_int), ("act", C.c_int), ("accumulate", C.c_int),
                ("tmode", C.c_int), ("barrier_before", C.c_int), ("bn_relu", C.c_int),
                ("x", _P), ("ldx", C.c_longlong), ("w", _P), ("bias", _P), ("out", _P), ("ldo", C.c_longlong),
                ("res", _P), ("ldr", C.c_longlong), ("gidx", _P), ("gtab", _P), ("ldt", C.c_longlong),
                ("bn_mean", _P), ("bn_var", _P), ("bn_gamma", _P), ("bn_beta", _P), ("bn_eps", C.c_float), ("ln_eps", C.c_float),
                ("t", _P), ("t_dev", _P)]


class ClusterOp(C.Structure):
    """One op of the cluster-resident eval forward (csrc/dense_cluster.cu ``ClusterOp``): activations are addressed by their
    offset in the per-CTA shared-memory arena, not by pointer."""
    _fields_ = [("kind", C.c_int), ("N", C.c_int), ("K", C.c_int), ("act", C.c_int), ("accumulate", C.c_int),
                ("tmode", C.c_int), ("bn_relu", C.c_int), ("out_global", C.c_int),
                ("x_off", C.c_int), ("ldx", C.c_int), ("out_off", C.c_int), ("ldo", C.c_int),
                ("res_off", C.c_int), ("ldr", C.c_int), ("fcp", C.c_int), ("kc", C.c_int),
                ("w", _P), ("bias", _P), ("tmap", _P), ("pad_", C.c_longlong),
                ("gx", _P), ("gldx", C.c_longlong), ("gout", _P), ("gldo", C.c_longlong),
                ("gidx", _P), ("gtab", _P), ("ldt", C.c_longlong),
                ("bn_mean", _P), ("bn_var", _P), ("bn_gamma", _P), ("bn_beta", _P), ("bn_eps", C.c_float), ("ln_eps", C.c_float),
                ("t", _P), ("t_dev", _P)]


FUSED_MAX_BATCH = 1024
CLUSTER_MAX_BATCH = 256               # the cluster kernel is tried up to here; it declines batches that need a second pass (> 135 rows)                # above this the per-layer GEMM kernels (tcgen05 tf32 at >= 2048) are the better shape


def cluster_arena_layout(recs: List[dict], ld: Dict[str, int], rows: int, arena_floats: int) -> Optional[Dict[str, int]]:
    """Offsets (floats) of the activation buffers in the cluster kernel's per-CTA shared-memory arena, first-fit by liveness.

    ``recs``: the tape in execution order, each ``{"kind", "reads": [Mat], "write": Mat, "global": bool}``; a kind-0 record
    that is not ``global`` is a Linear whose output is PUSHED into the peers' arenas and ends an *epoch* (the peers only pace
    each other through those pushes).  While a CTA pushes the output of the Linear that ends epoch e, a slower peer may still be
    anywhere inside epoch e: in the row-wise ops in front of that Linear or in its own FMA loop.  So a region is reused for
    an op's output only if its previous owner was last touched in an EARLIER epoch.  ``eps`` (written to global memory) gets no
    region.  Returns None when the arena is too small."""
    epoch, e = [], 0
    for r in recs:
        epoch.append(e)
        if r["kind"] == 0 and not r["global"]:
            e += 1
    last_use: Dict[str, int] = {}
    for j, r in enumerate(recs):
        for m in r["reads"] + [r["write"]]:
            last_use[m[0]] = j
    base: Dict[str, int] = {}
    live: List[Tuple[int, int, str]] = []                    # (offset, size, name)
    for j, r in enumerate(recs):
        name = r["write"][0]
        if name == "eps" or name in base:
            continue
        live = [b for b in live if epoch[last_use[b[2]]] >= epoch[j]]
        size, off = rows * ld[name], 0
        for b in sorted(live):
            if off + size <= b[0]:
                break
            off = max(off, b[0] + b[1])
        if off + size > arena_floats:
            return None
        base[name] = off
        live.append((off, size, name))
    return base


class DenseEngine:
    def __init__(self, module: torch.nn.Module, batch: int, device: torch.device, training: bool, in_dim: int,
                 emb_mode: int):
        self.module, self.B, self.device, self.training = module, batch, device, training
        self.lib = L.load()
        self.emb_mode = emb_mode
        self.widths: Dict[str, int] = {}
        self.bufs: Dict[str, torch.Tensor] = {}
        self.gbufs: Dict[str, torch.Tensor] = {}
        self._ops: List[dict] = []
        self.pgrad: Dict[str, torch.Tensor] = {k: torch.zeros_like(p, device=device) for k, p in module.named_parameters()}
        self._pname = {id(p): k for k, p in module.named_parameters()}
        self.x_in = self.new("x_in", in_dim)
        self.eps = self.new("eps", in_dim)
        self.d_eps = self.gbufs["eps"]
        self.t_in = torch.zeros(batch, device=device, dtype=torch.int64)
        self.t_dev = torch.zeros(1, device=device, dtype=torch.int32)
        self.y_in = torch.zeros(batch, device=device, dtype=torch.int64)
        self.zero_idx = torch.zeros(batch, device=device, dtype=torch.int64)
        self.use_t_dev = False
        self.seed = torch.zeros(2, device=device, dtype=torch.int64)       # dropout: (seed, unused)
        self._drop_slots: List[torch.Tensor] = []
        self.flops = 0.0
        self.cfg = type("Cfg", (), {"cond": "class"})()
        self.fwd_ops: List[Tuple[str, Callable[[int], None]]] = []
        self.bwd_ops: List[Tuple[str, Callable[[int], None]]] = []

    # ------------------------------------------------------------------ buffers
    def new(self, name: str, width: int) -> torch.Tensor:
        self.widths[name] = width
        self.bufs[name] = torch.zeros(self.B, width, device=self.device)
        self.gbufs[name] = torch.zeros(self.B, width, device=self.device)
        return self.bufs[name]

    def full(self, name: str) -> Mat:
        return (name, 0, self.widths[name])

    def val(self, m: Mat) -> torch.Tensor:
        return self.bufs[m[0]][:, m[1]:m[2]]

    def grad(self, m: Mat) -> torch.Tensor:
        return self.gbufs[m[0]][:, m[1]:m[2]]

    # ------------------------------------------------------------------ op declarations
    def time_features(self, out: Mat):
        self._ops.append({"kind": "time", "out": out})

    def linear(self, name: str, x: Mat, weight, bias, out: Mat, act: int = L.ACT_NONE, pre: Optional[Mat] = None,
               residual: Optional[Mat] = None, gather=None, rows: Optional[Tuple[int, int]] = None,
               x_needs_grad: bool = True):
        """out = act(x W[rows]^T + b[rows]) + residual + table[idx].  ``gather`` = (index tensor, Parameter)."""
        self._ops.append({"kind": "linear", "name": name, "x": x, "w": weight, "b": bias, "out": out, "act": act,
                          "pre": pre, "res": residual, "gather": gather, "rows": rows, "xg": x_needs_grad})

    def bn1d(self, name: str, x: Mat, bn: torch.nn.BatchNorm1d, out: Mat, relu: bool = True):
        self._ops.append({"kind": "bn", "name": name, "x": x, "bn": bn, "out": out, "relu": relu})

    def layernorm(self, name: str, x: Mat, ln: torch.nn.LayerNorm, out: Mat):
        self._ops.append({"kind": "ln", "name": name, "x": x, "ln": ln, "out": out})

    def dropout(self, name: str, x: Mat, out: Mat, p: float, group: int = 1):
        if self.training and p > 0.0:
            slot = torch.zeros(2, device=self.device, dtype=torch.int64)
            slot[1] = len(self._drop_slots) + 1
            self._drop_slots.append(slot)
            self._ops.append({"kind": "drop", "name": name, "x": x, "out": out, "p": float(p), "group": group,
                              "slot": slot})
        else:
            self._ops.append({"kind": "copy", "name": name, "x": x, "out": out, "acc": 0})

    def add_into(self, name: str, x: Mat, out: Mat, accumulate: bool):
        """out (+)= x   (two of these build a residual sum when dropout sits between GEMM and add)."""
        self._ops.append({"kind": "copy", "name": name, "x": x, "out": out, "acc": int(accumulate)})

    # ------------------------------------------------------------------ plan construction
    def _tscratch(self, which: int, numel: int) -> torch.Tensor:
        """One of the two scratch vectors that hold the transposed operands of a large-batch weight gradient."""
        return self._ts[which][:numel]

    def _colsum_part(self, numel: int) -> torch.Tensor:
        """Partial-sum workspace of the large-batch bias gradients (one buffer, the bias gradients run one after another)."""
        cur = getattr(self, "_cs_part", None)
        if cur is None or cur.numel() < numel:
            # earlier closures keep their (smaller) buffer alive; every step list captures the buffer it was built with
            cur = torch.zeros(numel, device=self.device)
            self._cs_part = cur
        return cur

    def _bn_streaming(self, op, x, N: int) -> bool:
        """Train-mode BatchNorm1d + ReLU of a large batch on the many-CTA NHWC BatchNorm kernels: needs a contiguous
        input and a channel count the 16-byte-vector kernels accept (4 * 2^k <= 1024)."""
        lanes = N // 4
        return (self.training and self.B >= 2048 and bool(op["relu"]) and x.stride(0) == N and N % 4 == 0
                and lanes & (lanes - 1) == 0 and lanes <= 256)

    def _splitk_ws(self, numel: int) -> torch.Tensor:
        """Split-K workspace of the batch-reducing weight-gradient GEMMs at large batch (one buffer: they run one after another)."""
        cur = getattr(self, "_sk_ws", None)
        if cur is None or cur.numel() < numel:
            cur = torch.empty(numel, device=self.device)
            self._sk_ws = cur
        return cur

    def _gemm(self, M, N, K, A, a_rs, a_cs, Bm, b_rs, b_cs, Cm, ldc, bias=None, act=0, pre=None, ld_pre=0, res=None,
              ldr=0, gidx=None, gtab=None, ldt=0, accumulate=0, splitk=False):
        g = L.GemmArgs()
        g.M, g.N, g.K, g.alpha = M, N, K, 1.0
        g.A, g.a_rs, g.a_cs = A, a_rs, a_cs
        g.B, g.b_rs, g.b_cs = Bm, b_rs, b_cs
        g.C, g.ldc, g.bias, g.act = Cm, ldc, bias, act
        g.pre_out, g.ld_pre, g.residual, g.ldr = pre, ld_pre, res, ldr
        g.gather_idx, g.gather_table, g.ld_table = gidx, gtab, ldt
        g.accumulate, g.splitk_ws = accumulate, None
        ws = None
        if splitk:                                  # K = the batch: slices over K, fixed-order second pass (FFMA and tcgen05 paths)
            need = int(self.lib.td_gemm_f32_workspace(M, N, K))
            if need > 0:
                ws = self._splitk_ws(need)
                g.splitk_ws = ws.data_ptr()
        g.allow_tf32 = int(self.B >= 2048)          # large batches: tcgen05 kind::tf32 GEMM (linear_tc.cu), tolerance 1e-2
        self.flops += 2.0 * M * N * K
        lib = self.lib
        return lambda st, g=g, ws=ws: L.check(lib.td_gemm_f32(C.byref(g), st), "td_gemm_f32")

    def build(self):
        """Materialise the forward launches and derive the backward plan (reverse order)."""
        B, lib = self.B, self.lib
        fwd: List[Tuple[str, Callable]] = []
        self.flops = 0.0
        self._saved: Dict[str, Dict[str, torch.Tensor]] = {}
        for op in self._ops:
            k = op["kind"]
            if k == "time":
                out = self.val(op["out"])
                assert out.stride(0) == out.shape[1]
                mode, D = self.emb_mode, out.shape[1]

                def f(st, out=out, mode=mode, D=D):
                    t = None if self.use_t_dev else self.t_in.data_ptr()
                    L.check(lib.td_time_features(t, self.t_dev.data_ptr(), out.data_ptr(), B, D, mode, st),
                            "td_time_features")
                fwd.append(("time_features", f))
            elif k == "linear":
                x, out = self.val(op["x"]), self.val(op["out"])
                w, b = op["w"], op["b"]
                r0, r1 = op["rows"] if op["rows"] else (0, w.shape[0])
                K, N = w.shape[1], r1 - r0
                assert x.shape[1] == K and out.shape[1] == N, (op["name"], x.shape, w.shape, out.shape)
                wp = w.data_ptr() + 4 * r0 * K
                bp = (b.data_ptr() + 4 * r0) if b is not None else None
                pre = self.val(op["pre"]) if op["pre"] else None
                res = self.val(op["res"]) if op["res"] else None
                gi, gt = (op["gather"][0], op["gather"][1]) if op["gather"] else (None, None)
                fwd.append((op["name"], self._gemm(
                    B, N, K, x.data_ptr(), x.stride(0), 1, wp, 1, K, out.data_ptr(), out.stride(0), bp, op["act"],
                    L.ptr(pre), pre.stride(0) if pre is not None else 0, L.ptr(res),
                    res.stride(0) if res is not None else 0, L.ptr(gi), L.ptr(gt), gt.shape[-1] if gt is not None else 0)))
            elif k == "bn":
                x, out, bn = self.val(op["x"]), self.val(op["out"]), op["bn"]
                N = x.shape[1]
                sv = {"mean": torch.zeros(N, device=self.device), "rstd": torch.zeros(N, device=self.device)}
                self._saved[op["name"]] = sv
                tr = int(self.training)
                mom = float(bn.momentum if bn.momentum is not None else 0.1)

                if self._bn_streaming(op, x, N):
                    # large batch, train mode: the conv engine's BatchNorm kernels on the [B][N] matrix (many-CTA partial
                    # sums -> finalize -> apply) instead of td_bn1d_fwd, whose one CTA per 32 columns walks the whole batch
                    rows = int(lib.td_chan_reduce_rows(L.TD_F32, B, N))
                    sv.update({"scale": torch.zeros(N, device=self.device), "shift": torch.zeros(N, device=self.device),
                               "coef": torch.zeros(3, N, device=self.device),
                               "part": torch.zeros((rows * 2 + 1) * N, device=self.device), "rows": rows})

                    def fs(st, x=x, out=out, bn=bn, sv=sv, N=N, mom=mom, rows=rows):
                        L.check(lib.td_bn_stats(x.data_ptr(), L.TD_F32, N, 0, B, N, sv["part"].data_ptr(), 1, st), "td_bn_stats")
                        L.check(lib.td_bn_finalize(sv["part"].data_ptr(), rows, N, B, bn.weight.data_ptr(), bn.bias.data_ptr(),
                                                   None, float(bn.eps), mom, bn.running_mean.data_ptr(),
                                                   bn.running_var.data_ptr(), bn.num_batches_tracked.data_ptr(),
                                                   sv["scale"].data_ptr(), sv["shift"].data_ptr(), sv["mean"].data_ptr(),
                                                   sv["rstd"].data_ptr(), st), "td_bn_finalize")
                        L.check(lib.td_bn_relu_apply(x.data_ptr(), sv["scale"].data_ptr(), sv["shift"].data_ptr(),
                                                     out.data_ptr(), L.TD_F32, out.stride(0), 0, B, N, 1, st), "td_bn_relu_apply")
                    fwd.append((op["name"], fs))
                    continue

                def f(st, x=x, out=out, bn=bn, sv=sv, N=N, tr=tr, mom=mom, relu=int(op["relu"])):
                    L.check(lib.td_bn1d_fwd(x.data_ptr(), x.stride(0), bn.weight.data_ptr(), bn.bias.data_ptr(),
                                            bn.running_mean.data_ptr(), bn.running_var.data_ptr(),
                                            sv["mean"].data_ptr(), sv["rstd"].data_ptr(), out.data_ptr(), out.stride(0),
                                            B, N, float(bn.eps), mom, tr, relu, st), "td_bn1d_fwd")
                    if tr:
                        bn.num_batches_tracked.add_(1)
                fwd.append((op["name"], f))
            elif k == "ln":
                x, out, ln = self.val(op["x"]), self.val(op["out"]), op["ln"]
                D = x.shape[1]
                assert x.stride(0) == D and out.stride(0) == D
                sv = {"mean": torch.zeros(B, device=self.device), "rstd": torch.zeros(B, device=self.device)}
                self._saved[op["name"]] = sv
                fwd.append((op["name"], lambda st, x=x, out=out, ln=ln, sv=sv, D=D: L.check(
                    lib.td_layernorm_fwd(x.data_ptr(), ln.weight.data_ptr(), ln.bias.data_ptr(), out.data_ptr(),
                                         sv["mean"].data_ptr(), sv["rstd"].data_ptr(), B, D, float(ln.eps), st),
                    "td_layernorm_fwd")))
            elif k == "drop":
                x, out, slot = self.val(op["x"]), self.val(op["out"]), op["slot"]
                fwd.append((op["name"], lambda st, x=x, out=out, slot=slot, p=op["p"], gsz=op["group"]: L.check(
                    lib.td_dropout_f32(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), B, x.shape[1], gsz, p,
                                       slot.data_ptr(), st), "td_dropout_f32")))
            elif k == "copy":
                x, out = self.val(op["x"]), self.val(op["out"])
                if x.data_ptr() != out.data_ptr():
                    fwd.append((op["name"], lambda st, x=x, out=out, acc=op["acc"]: L.check(
                        lib.td_add2d_f32(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), B, x.shape[1], acc, st),
                        "td_add2d_f32")))
        self.fwd_ops = fwd
        self.fwd_flops = self.flops
        self.bwd_ops = self._build_backward() if self.training else []
        self._tapes = None
        self._ctapes = None
        mode = os.environ.get("TD_DENSE_FUSED", "1")       # 0: one launch per op, 1: cluster kernel / tape kernel, 2: tape kernel only
        if not self.training and mode != "0":
            if self.B <= CLUSTER_MAX_BATCH and mode != "2":
                self._build_cluster_tapes()
            if (self._ctapes is None and self.B <= FUSED_MAX_BATCH
                    and all(op["w"].shape[1] <= 1024 for op in self._ops if op["kind"] == "linear")):      # K staged whole in smem
                self._build_tapes()

    # ------------------------------------------------------------------ fused eval-mode forward (one persistent kernel)
    def _build_tapes(self):
        """Compile the declared ops into the device tape of csrc/dense_fused.cu: eval-mode BatchNorm1d (+ReLU) folded into the
        Linear that feeds it, dropout / copies dropped or turned into adds, and a grid barrier in front of every op that
        touches a buffer written since the previous barrier.  Two tapes: per-sample ``t_in`` and the sampler's ``t_dev``."""
        assert int(self.lib.td_dense_tape_op_bytes()) == C.sizeof(TapeOp), "TapeOp layout mismatch"
        ops = list(self._ops)
        fused_bn = {}
        for i, op in enumerate(ops):                      # linear -> bn on exactly the linear's output buffer
            if op["kind"] == "bn" and i > 0 and ops[i - 1]["kind"] == "linear" and ops[i - 1]["out"] == op["x"] \
                    and ops[i - 1]["act"] == L.ACT_NONE and ops[i - 1]["res"] is None and ops[i - 1]["gather"] is None:
                fused_bn[i - 1] = op
        skip = {id(v) for v in fused_bn.values()}

        def make(use_t_dev: bool):
            tape, dirty, seen = [], set(), set()          # buffers written / read since the last barrier

            def deps(reads, writes):
                reads = [r for r in reads if r is not None]
                hit = (any(r[0] in dirty for r in reads) or any(w[0] in dirty for w in writes)      # RAW / WAW
                       or any(w[0] in seen for w in writes))                                        # WAR
                if hit:
                    dirty.clear()
                    seen.clear()
                dirty.update(w[0] for w in writes)
                seen.update(r[0] for r in reads)
                return int(hit)

            for i, op in enumerate(ops):
                if id(op) in skip:
                    continue
                k = op["kind"]
                t = TapeOp()
                if k == "time":
                    out = self.val(op["out"])
                    t.kind, t.N, t.tmode = 3, out.shape[1], self.emb_mode
                    t.out, t.ldo = out.data_ptr(), out.stride(0)
                    t.t = None if use_t_dev else self.t_in.data_ptr()
                    t.t_dev = self.t_dev.data_ptr()
                    t.barrier_before = deps([], [op["out"]])
                elif k == "linear":
                    x, w, b = self.val(op["x"]), op["w"], op["b"]
                    r0, r1 = op["rows"] if op["rows"] else (0, w.shape[0])
                    dst = op["out"]
                    bn = fused_bn.get(i)
                    if bn is not None:
                        dst = bn["out"]
                        m = bn["bn"]
                        t.bn_mean, t.bn_var = m.running_mean.data_ptr(), m.running_var.data_ptr()
                        t.bn_gamma, t.bn_beta = m.weight.data_ptr(), m.bias.data_ptr()
                        t.bn_eps, t.bn_relu = float(m.eps), int(bn["relu"])
                    out = self.val(dst)
                    t.kind, t.N, t.K, t.act = 0, r1 - r0, w.shape[1], op["act"]
                    t.x, t.ldx = x.data_ptr(), x.stride(0)
                    t.w = w.data_ptr() + 4 * r0 * w.shape[1]
                    t.bias = (b.data_ptr() + 4 * r0) if b is not None else None
                    t.out, t.ldo = out.data_ptr(), out.stride(0)
                    if op["res"] is not None:
                        r = self.val(op["res"])
                        t.res, t.ldr = r.data_ptr(), r.stride(0)
                    if op["gather"] is not None:
                        gi, gt = op["gather"]
                        t.gidx, t.gtab, t.ldt = gi.data_ptr(), gt.data_ptr(), gt.shape[-1]
                    t.barrier_before = deps([op["x"], op["res"]], [dst])
                elif k == "bn":                            # a BatchNorm1d that does not follow its Linear directly: not declared by either model
                    raise NotImplementedError("stand-alone BatchNorm1d in the fused tape")
                elif k == "ln":
                    x, out, ln = self.val(op["x"]), self.val(op["out"]), op["ln"]
                    t.kind, t.N = 1, x.shape[1]
                    t.x, t.ldx, t.out, t.ldo = x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0)
                    t.w, t.bias, t.ln_eps = ln.weight.data_ptr(), ln.bias.data_ptr(), float(ln.eps)
                    t.barrier_before = deps([op["x"]], [op["out"]])
                elif k in ("copy", "drop"):                # eval mode: dropout is the identity
                    x, out = self.val(op["x"]), self.val(op["out"])
                    if x.data_ptr() == out.data_ptr():
                        continue
                    t.kind, t.N, t.accumulate = 2, x.shape[1], int(op.get("acc", 0))
                    t.x, t.ldx, t.out, t.ldo = x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0)
                    t.barrier_before = deps([op["x"]], [op["out"]])
                else:
                    raise NotImplementedError(k)
                tape.append(t)
            raw = b"".join(bytes(t) for t in tape)
            return torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.device), len(tape), sum(t.barrier_before for t in tape)

        self._tape_bar = torch.zeros(1, device=self.device, dtype=torch.int64)
        self._tapes = {False: make(False), True: make(True)}

    # ------------------------------------------------------------------ cluster-resident eval forward (csrc/dense_cluster.cu)
    def _build_cluster_tapes(self):
        """Compile the declared ops for ``dense_cluster_kernel``: every activation buffer gets an offset in the per-CTA
        shared-memory arena (first-fit by liveness), buffers the tape reads but never writes are loaded from global memory
        first, and the op that writes ``eps`` stores to global memory.  A Linear ends with a cluster barrier (its output
        is pushed into the peers' arenas); an arena region is only reused for an op's output when its previous owner was
        last touched BEFORE the preceding barrier -- a fast CTA may push a Linear's output while a slow peer still runs
        the row-wise ops in front of that Linear.  Leaves ``_ctapes`` None when the model does not fit (the grid-barrier
        tape then takes over)."""
        lib = self.lib
        assert int(lib.td_dense_cluster_op_bytes()) == C.sizeof(ClusterOp), "ClusterOp layout mismatch"
        lim = [C.c_int() for _ in range(6)]
        L.check(lib.td_dense_cluster_limits(self.B, *[C.byref(v) for v in lim]), "td_dense_cluster_limits")
        R, CL, ARENA, STAGE, MAX_OPS, MAX_CLUSTERS = (int(v.value) for v in lim)
        if -(-self.B // R) > MAX_CLUSTERS and os.environ.get("TD_DENSE_CLUSTER_PASSES", "1") == "1":
            return              # a second pass re-streams every weight: measured no faster than the grid-barrier tape (batch 200)
        ops = list(self._ops)
        fused_bn = {}
        for i, op in enumerate(ops):
            if op["kind"] == "bn" and i > 0 and ops[i - 1]["kind"] == "linear" and ops[i - 1]["out"] == op["x"] \
                    and ops[i - 1]["act"] == L.ACT_NONE and ops[i - 1]["res"] is None and ops[i - 1]["gather"] is None:
                fused_bn[i - 1] = op
        skip = {id(v) for v in fused_bn.values()}
        # abstract records: (kind, reads, write, payload)
        recs = []
        for i, op in enumerate(ops):
            if id(op) in skip:
                continue
            k = op["kind"]
            if k == "time":
                recs.append({"kind": 3, "reads": [], "write": op["out"], "op": op})
            elif k == "linear":
                bn = fused_bn.get(i)
                dst = bn["out"] if bn is not None else op["out"]
                recs.append({"kind": 0, "reads": [m for m in (op["x"], op["res"]) if m is not None], "write": dst, "op": op, "bn": bn})
            elif k == "ln":
                recs.append({"kind": 1, "reads": [op["x"]], "write": op["out"], "op": op})
            elif k in ("copy", "drop"):
                if op["x"] == op["out"]:
                    continue
                reads = [op["x"]] + ([op["out"]] if op.get("acc", 0) else [])
                recs.append({"kind": 2, "reads": reads, "write": op["out"], "op": op})
            else:                                           # stand-alone BatchNorm1d: not declared by either model
                return
        # Two Linears back to back with nothing in between (the length-1 attention's out_proj(v_proj(x)), diffusion_transformer.py:99)
        # are ONE Linear with W = W_b W_a, b = W_b b_a + b_b in eval mode: one op, one push and one barrier wait fewer.  The merged
        # weights are derived buffers, recomputed by refresh_weights() when a parameter changes.
        self._merged = []
        if os.environ.get("TD_DENSE_MERGE", "1") != "0":
            n_reads: Dict[str, int] = {}
            for r in recs:
                for m in r["reads"]:
                    n_reads[m[0]] = n_reads.get(m[0], 0) + 1
            out, i = [], 0
            while i < len(recs):
                a = recs[i]
                b = recs[i + 1] if i + 1 < len(recs) else None
                ok = (b is not None and a["kind"] == 0 and b["kind"] == 0 and a["bn"] is None and a["op"]["act"] == L.ACT_NONE
                      and a["op"]["res"] is None and a["op"]["gather"] is None and a["write"] == self.full(a["write"][0])
                      and b["op"]["x"] == a["write"] and n_reads.get(a["write"][0], 0) == 1 and a["write"][0] != "eps")
                if not ok:
                    out.append(a)
                    i += 1
                    continue
                wa, ba, wb, bb = a["op"]["w"], a["op"]["b"], b["op"]["w"], b["op"]["b"]
                ra = a["op"]["rows"] if a["op"]["rows"] else (0, wa.shape[0])
                rb = b["op"]["rows"] if b["op"]["rows"] else (0, wb.shape[0])
                mg = {"wa": wa, "ba": ba, "ra": ra, "wb": wb, "bb": bb, "rb": rb,
                      "W": torch.zeros(rb[1] - rb[0], wa.shape[1], device=self.device),
                      "b": torch.zeros(rb[1] - rb[0], device=self.device)}
                self._merged.append(mg)
                out.append({"kind": 0, "reads": [a["op"]["x"]] + ([b["op"]["res"]] if b["op"]["res"] is not None else []),
                            "write": b["write"], "op": b["op"], "bn": b["bn"], "merged": mg, "x": a["op"]["x"]})
                i += 2
            recs = out
        # buffers read before any op wrote them come from global memory
        written, loads = set(), []
        for r in recs:
            for m in r["reads"]:
                if m[0] not in written and m[0] not in [l["write"][0] for l in loads]:
                    loads.append({"kind": 4, "reads": [], "write": self.full(m[0]), "op": None})
            written.add(r["write"][0])
        recs = loads + recs
        if len(recs) > MAX_OPS:
            return
        for j, r in enumerate(recs):
            r["global"] = r["kind"] == 0 and r["write"][0] == "eps"
            if r["write"][0] == "eps" and (not r["global"] or j != len(recs) - 1 or r["write"] != self.full("eps")):
                return
        if any(m[0] == "eps" for r in recs for m in r["reads"]) or not recs[-1]["global"]:
            return
        ld = {n: (w + 3) // 4 * 4 for n, w in self.widths.items()}
        base = cluster_arena_layout(recs, ld, R, ARENA)
        if base is None:
            return

        # one tensor map per streamed weight (TMA boxes of 32 floats x fcp rows), kept on the device beside the tape
        n_lin = sum(1 for r in recs if r["kind"] == 0)
        maps_host = torch.zeros(max(n_lin, 1) * 128, dtype=torch.uint8)
        maps_dev = torch.zeros(max(n_lin, 1) * 128, dtype=torch.uint8, device=self.device)
        map_slot: Dict[int, int] = {}

        def make(use_t_dev: bool):
            tape = []
            for r in recs:
                t = ClusterOp()
                t.kind, t.res_off = r["kind"], -1
                wname, c0, c1 = r["write"]
                if r["global"]:
                    out = self.val(r["write"])
                    t.out_global, t.gout, t.gldo = 1, out.data_ptr(), out.stride(0)
                else:
                    t.out_off, t.ldo = base[wname] + c0, ld[wname]
                op = r["op"]
                if r["kind"] == 4:
                    src = self.bufs[wname]
                    t.N, t.gx, t.gldx = c1 - c0, src.data_ptr(), src.stride(0)
                elif r["kind"] == 3:
                    t.N, t.tmode = c1 - c0, self.emb_mode
                    t.t = None if use_t_dev else self.t_in.data_ptr()
                    t.t_dev = self.t_dev.data_ptr()
                elif r["kind"] == 0:
                    mg = r.get("merged")
                    if mg is not None:
                        N, K = mg["W"].shape
                        xn, x0, x1 = r["x"]
                        t.w, t.bias = mg["W"].data_ptr(), mg["b"].data_ptr()
                    else:
                        w, b = op["w"], op["b"]
                        r0, r1 = op["rows"] if op["rows"] else (0, w.shape[0])
                        N, K = r1 - r0, w.shape[1]
                        xn, x0, x1 = op["x"]
                        t.w = w.data_ptr() + 4 * r0 * K
                        t.bias = (b.data_ptr() + 4 * r0) if b is not None else None
                    assert x1 - x0 == K and c1 - c0 == N
                    t.N, t.K, t.act = N, K, op["act"]
                    t.x_off, t.ldx = base[xn] + x0, ld[xn]
                    fcp = 4
                    while fcp * CL < N:
                        fcp *= 2
                    if fcp > 128:
                        return None
                    if not r["global"] and (CL - 1) * fcp >= N:
                        return None          # a CTA without features of a pushed op would not be paced by its peers (no cluster barrier)
                    t.fcp = fcp
                    kc = 32 * (STAGE // (fcp * 32))                  # whole boxes per stage
                    t.kc = kc if (K % 4 == 0 and t.w % 16 == 0 and t.x_off % 4 == 0 and kc >= 32) else 0
                    if t.kc:
                        slot = map_slot.setdefault(id(r), len(map_slot))
                        if L.load().td_dense_cluster_weight_map(t.w, N, K, fcp, maps_host.data_ptr() + 128 * slot) != 0:
                            return None                              # no driver entry point: the tape kernel takes over
                        t.tmap = maps_dev.data_ptr() + 128 * slot
                    if r["bn"] is not None:
                        m = r["bn"]["bn"]
                        t.bn_mean, t.bn_var = m.running_mean.data_ptr(), m.running_var.data_ptr()
                        t.bn_gamma, t.bn_beta = m.weight.data_ptr(), m.bias.data_ptr()
                        t.bn_eps, t.bn_relu = float(m.eps), int(r["bn"]["relu"])
                    if op["res"] is not None:
                        rn, q0, _ = op["res"]
                        t.res_off, t.ldr = base[rn] + q0, ld[rn]
                    if op["gather"] is not None:
                        gi, gt = op["gather"]
                        t.gidx, t.gtab, t.ldt = gi.data_ptr(), gt.data_ptr(), gt.shape[-1]
                elif r["kind"] == 1:
                    ln = op["ln"]
                    xn, x0, x1 = op["x"]
                    t.N, t.x_off, t.ldx = x1 - x0, base[xn] + x0, ld[xn]
                    t.w, t.bias, t.ln_eps = ln.weight.data_ptr(), ln.bias.data_ptr(), float(ln.eps)
                else:
                    xn, x0, x1 = op["x"]
                    t.N, t.accumulate = x1 - x0, int(op.get("acc", 0))
                    t.x_off, t.ldx = base[xn] + x0, ld[xn]
                tape.append(t)
            raw = b"".join(bytes(t) for t in tape)
            return torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.device), len(tape)

        tapes = {False: make(False), True: make(True)}
        if tapes[False] is None or tapes[True] is None:
            return
        maps_dev.copy_(maps_host)
        self._ctape_maps = maps_dev
        self._ctapes = tapes
        self._merged_version = None
        self.refresh_weights()
        self._ctape_rows = R
        self._ctape_arena = max(b + R * ld[n] for n, b in base.items())

    def can_fuse_step(self) -> bool:
        """The sampler's reverse-step update can ride in the cluster kernel's last epilogue (one launch per reverse step)."""
        return (getattr(self, "_ctapes", None) is not None and os.environ.get("TD_DENSE_STEP_FUSED", "1") != "0"
                and self.widths["x_in"] % 4 == 0 and bool(self.lib.td_dense_cluster_step_fused()))

    def fused_step(self, z_ptr, z_stride, seed_ptr, coef_ptr, num_timesteps, ticket_ptr) -> None:
        """eps = model(x_in, t_dev, y) and x_in <- c1 (x_in - c2 eps) + c3 z, t_dev -= 1, in one launch (``td_dense_cluster_step``)."""
        buf, n = self._ctapes[True]
        L.check(self.lib.td_dense_cluster_step(buf.data_ptr(), n, self.B, self._ctape_rows, 0, self.x_in.data_ptr(), coef_ptr, z_ptr,
                                               z_stride, seed_ptr, self.t_dev.data_ptr(), ticket_ptr, num_timesteps, L.stream_ptr()),
                "td_dense_cluster_step")

    def _launch_tape(self, st: int) -> None:
        if self._ctapes is not None:
            buf, n = self._ctapes[bool(self.use_t_dev)]
            L.check(self.lib.td_dense_cluster_run(buf.data_ptr(), n, self.B, self._ctape_rows, 0, st), "td_dense_cluster_run")
            return
        buf, n, _ = self._tapes[bool(self.use_t_dev)]
        L.check(self.lib.td_dense_tape_run(buf.data_ptr(), n, self.B, self._tape_bar.data_ptr(), 0, st), "td_dense_tape_run")

    def _build_backward(self):
        B, lib = self.B, self.lib
        bwd: List[Tuple[str, Callable]] = []
        written: Dict[str, List[Tuple[int, int]]] = {}

        def acc_flag(m: Mat) -> int:
            """0 = first gradient to reach these columns (overwrite), 1 = accumulate."""
            spans = written.setdefault(m[0], [])
            for a, b in spans:
                if a <= m[1] and m[2] <= b:                 # already covered by an earlier (wider or equal) write
                    return 1
                assert b <= m[1] or a >= m[2], f"partially overlapping gradient slices on {m[0]}"
            spans.append((m[1], m[2]))
            return 0

        scratch_w = max(self.widths.values())
        self._scratch = torch.zeros(self.B, scratch_w, device=self.device)
        self._wt: List[Tuple[torch.Tensor, torch.Tensor, int, int]] = []
        self._ts = [torch.empty(self.B * scratch_w, device=self.device) for _ in range(2)] if self.B >= 2048 else None
        written["eps"] = [(0, self.widths["eps"])]               # seeded by the loss
        for op in reversed(self._ops):
            k = op["kind"]
            if k == "time":
                continue
            if k == "linear":
                x, out, w, b = self.val(op["x"]), self.val(op["out"]), op["w"], op["b"]
                g_out = self.grad(op["out"])
                r0, r1 = op["rows"] if op["rows"] else (0, w.shape[0])
                K, N = w.shape[1], r1 - r0
                name = op["name"]
                steps: List[Callable] = []
                if op["res"] is not None:                       # d residual (+)= d out (post-activation gradient)
                    gr = self.grad(op["res"])
                    af = acc_flag(op["res"])
                    steps.append(lambda st, s=g_out, d=gr, af=af: L.check(
                        lib.td_add2d_f32(s.data_ptr(), s.stride(0), d.data_ptr(), d.stride(0), B, s.shape[1], af, st),
                        "td_add2d_f32"))
                if op["gather"] is not None:
                    idx, tab = op["gather"]
                    tg = self.pgrad[self._pname[id(tab)]]
                    rows_t = tg.numel() // tg.shape[-1]
                    ews = self._colsum_part(min(296, max(2, B // 256)) * rows_t * tg.shape[-1]) if B >= 4096 else None
                    steps.append(lambda st, s=g_out, idx=idx, tg=tg, ews=ews, rows_t=rows_t: L.check(
                        lib.td_embedding_bwd(s.data_ptr(), s.stride(0), idx.data_ptr(), tg.data_ptr(), B, tg.shape[-1],
                                             rows_t, 0, L.ptr(ews), ews.numel() if ews is not None else 0, st), "td_embedding_bwd"))
                g_pre = g_out
                if op["act"] != L.ACT_NONE:
                    pre = self.val(op["pre"])
                    assert g_out.stride(0) == N and pre.stride(0) == N, "activated outputs must be whole buffers"
                    g_pre = self._scratch.view(-1)[:B * N].view(B, N)
                    steps.append(lambda st, s=g_out, pre=pre, d=g_pre, act=op["act"], n=B * N: L.check(
                        lib.td_act_bwd_f32(s.data_ptr(), pre.data_ptr(), d.data_ptr(), n, act, st), "td_act_bwd_f32"))
                wg = self.pgrad[self._pname[id(w)]]
                wgp = wg.data_ptr() + 4 * r0 * K
                if B >= 2048 and B % 4 == 0 and K >= 32:
                    # large batch: dW = g^T x on the tensor-core GEMM.  The reduction runs over the batch and kind::tf32
                    # wants it contiguous in both operands: transpose g and x into two scratch matrices first
                    gt = self._tscratch(0, N * B).view(N, B)
                    xt = self._tscratch(1, K * B).view(K, B)
                    steps.append(lambda st, g=g_pre, x=x, gt=gt, xt=xt, N=N, K=K: (
                        L.check(lib.td_transpose_f32(g.data_ptr(), g.stride(0), gt.data_ptr(), B, B, N, st), "td_transpose_f32"),
                        L.check(lib.td_transpose_f32(x.data_ptr(), x.stride(0), xt.data_ptr(), B, B, K, st), "td_transpose_f32")))
                    steps.append(self._gemm(N, K, B, gt.data_ptr(), B, 1, xt.data_ptr(), 1, B, wgp, K, splitk=True))
                else:
                    steps.append(self._gemm(N, K, B, g_pre.data_ptr(), 1, g_pre.stride(0), x.data_ptr(), x.stride(0), 1,
                                            wgp, K, splitk=B >= 2048))                 # dW = g^T x
                if b is not None:
                    bg = self.pgrad[self._pname[id(b)]]
                    bgp = bg.data_ptr() + 4 * r0
                    lanes = N // 4
                    if (B >= 2048 and N % 4 == 0 and lanes & (lanes - 1) == 0 and lanes <= 256 and g_pre.stride(0) % 4 == 0
                            and g_pre.data_ptr() % 16 == 0):
                        # large batch: many-CTA partial column sums + fixed-order finalize (the conv engine's kernels)
                        # instead of td_colsum_f32, whose one CTA per 32 columns walks the whole batch
                        rows = int(lib.td_chan_reduce_rows(L.TD_F32, B, N))
                        part = self._colsum_part((rows * 2 + 1) * N)
                        steps.append(lambda st, s=g_pre, bgp=bgp, n=N, rows=rows, part=part: (
                            L.check(lib.td_bn_stats(s.data_ptr(), L.TD_F32, s.stride(0), 0, B, n, part.data_ptr(), 0, st), "td_bn_stats"),
                            L.check(lib.td_partial_sum(part.data_ptr(), rows, n, 0, bgp, st), "td_partial_sum")))
                    else:
                        steps.append(lambda st, s=g_pre, bgp=bgp, n=N: L.check(
                            lib.td_colsum_f32(s.data_ptr(), s.stride(0), bgp, B, n, 0, st), "td_colsum_f32"))
                if op["xg"]:
                    gx = self.grad(op["x"])
                    af = acc_flag(op["x"])
                    if B >= 2048:
                        # large batch: dx (+)= g W through the tensor-core GEMM, which wants both operands K-major
                        # (kind::tf32, linear_tc.cu): a transposed copy of the weight rows, refreshed once per backward
                        wt = torch.empty(K, N, device=self.device)
                        self._wt.append((wt, w, r0, r1))
                        steps.append(self._gemm(B, K, N, g_pre.data_ptr(), g_pre.stride(0), 1, wt.data_ptr(), 1, N,
                                                gx.data_ptr(), gx.stride(0), accumulate=af))
                    else:
                        wp = w.data_ptr() + 4 * r0 * K
                        steps.append(self._gemm(B, K, N, g_pre.data_ptr(), g_pre.stride(0), 1, wp, K, 1, gx.data_ptr(),
                                                gx.stride(0), accumulate=af))         # dx (+)= g W
                bwd.append((f"{name}:bwd", lambda st, steps=steps: [s(st) for s in steps] and None))
            elif k == "bn":
                x, out, bn = self.val(op["x"]), self.val(op["out"]), op["bn"]
                g_out, gx = self.grad(op["out"]), self.grad(op["x"])
                assert acc_flag(op["x"]) == 0, "BatchNorm1d input gradient must be the first writer"
                sv = self._saved[op["name"]]
                dg = self.pgrad[self._pname[id(bn.weight)]]
                db = self.pgrad[self._pname[id(bn.bias)]]
                N = x.shape[1]
                if "part" in sv:                    # large batch: the conv engine's BatchNorm backward kernels (see the forward)
                    assert gx.stride(0) == N

                    def bs(st, x=x, g_out=g_out, gx=gx, sv=sv, dg=dg, db=db, N=N):
                        L.check(lib.td_bn_relu_bwd_reduce(g_out.data_ptr(), g_out.stride(0), 0, x.data_ptr(), L.TD_F32,
                                                          sv["scale"].data_ptr(), sv["shift"].data_ptr(), sv["mean"].data_ptr(),
                                                          B, N, sv["part"].data_ptr(), st), "td_bn_relu_bwd_reduce")
                        L.check(lib.td_bn_bwd_finalize(sv["part"].data_ptr(), sv["rows"], N, B, sv["scale"].data_ptr(),
                                                       sv["mean"].data_ptr(), sv["rstd"].data_ptr(), dg.data_ptr(),
                                                       db.data_ptr(), sv["coef"].data_ptr(), st), "td_bn_bwd_finalize")
                        L.check(lib.td_bn_relu_bwd_apply(g_out.data_ptr(), g_out.stride(0), 0, x.data_ptr(), L.TD_F32,
                                                         sv["scale"].data_ptr(), sv["shift"].data_ptr(), sv["coef"].data_ptr(),
                                                         gx.data_ptr(), B, N, st), "td_bn_relu_bwd_apply")
                    bwd.append((f"{op['name']}:bwd", bs))
                    continue
                bwd.append((f"{op['name']}:bwd", lambda st, x=x, out=out, bn=bn, g_out=g_out, gx=gx, sv=sv, dg=dg, db=db,
                            N=N, relu=int(op["relu"]): L.check(
                    lib.td_bn1d_bwd(g_out.data_ptr(), g_out.stride(0), x.data_ptr(), x.stride(0), out.data_ptr(),
                                    out.stride(0), bn.weight.data_ptr(), sv["mean"].data_ptr(), sv["rstd"].data_ptr(),
                                    gx.data_ptr(), gx.stride(0), dg.data_ptr(), db.data_ptr(), B, N, relu, st),
                    "td_bn1d_bwd")))
            elif k == "ln":
                x, ln = self.val(op["x"]), op["ln"]
                g_out, gx = self.grad(op["out"]), self.grad(op["x"])
                assert acc_flag(op["x"]) == 0, "LayerNorm input gradient must be the first writer"
                sv = self._saved[op["name"]]
                dg = self.pgrad[self._pname[id(ln.weight)]]
                db = self.pgrad[self._pname[id(ln.bias)]]
                D = x.shape[1]
                assert g_out.stride(0) == D and gx.stride(0) == D
                lws = self._colsum_part(2 * D * min(296, max(2, B // 64))) if B >= 4096 else None
                bwd.append((f"{op['name']}:bwd", lambda st, x=x, ln=ln, g_out=g_out, gx=gx, sv=sv, dg=dg, db=db, D=D, lws=lws: L.check(
                    lib.td_layernorm_bwd(g_out.data_ptr(), x.data_ptr(), ln.weight.data_ptr(), sv["mean"].data_ptr(),
                                         sv["rstd"].data_ptr(), gx.data_ptr(), dg.data_ptr(), db.data_ptr(), B, D, L.ptr(lws),
                                         lws.numel() if lws is not None else 0, st),
                    "td_layernorm_bwd")))
            elif k == "drop":
                g_out, gx, slot = self.grad(op["out"]), self.grad(op["x"]), op["slot"]
                assert acc_flag(op["x"]) == 0
                bwd.append((f"{op['name']}:bwd", lambda st, g_out=g_out, gx=gx, slot=slot, p=op["p"], gsz=op["group"]: L.check(
                    lib.td_dropout_f32(g_out.data_ptr(), g_out.stride(0), gx.data_ptr(), gx.stride(0), B, g_out.shape[1],
                                       gsz, p, slot.data_ptr(), st), "td_dropout_f32")))
            elif k == "copy":
                g_out, gx = self.grad(op["out"]), self.grad(op["x"])
                if g_out.data_ptr() != gx.data_ptr():
                    af = acc_flag(op["x"])
                    bwd.append((f"{op['name']}:bwd", lambda st, s=g_out, d=gx, af=af: L.check(
                        lib.td_add2d_f32(s.data_ptr(), s.stride(0), d.data_ptr(), d.stride(0), B, s.shape[1], af, st),
                        "td_add2d_f32")))
        if self._wt:
            def refresh_wt(st):
                with torch.no_grad():
                    for wt, w, r0, r1 in self._wt:
                        wt.copy_(w.detach()[r0:r1].t())
            bwd.insert(0, ("weights^T", refresh_wt))
        return bwd

    # ------------------------------------------------------------------ UNetTrainEngine-compatible surface
    def _build(self):
        self.build()

    def weights_version(self):
        # _weights_gen: bumped by train.TrainStep, whose kernels update the parameters through raw pointers (no _version bump)
        return (getattr(self.module, "_weights_gen", 0),) + tuple(p._version for p in self.module.parameters())

    def refresh_weights(self, force: bool = False) -> None:
        """Parameters are read in place; only the merged weights of fused Linear pairs (cluster tape) are derived buffers."""
        merged = getattr(self, "_merged", None)
        if not merged or getattr(self, "_ctapes", None) is None:
            return
        ver = self.weights_version()
        if not force and ver == self._merged_version:
            return
        st, flops = L.stream_ptr(), self.flops
        for mg in merged:
            wa, wb = mg["wa"], mg["wb"]
            (a0, a1), (b0, b1) = mg["ra"], mg["rb"]
            Na, Ka, Nb = a1 - a0, wa.shape[1], b1 - b0
            assert wb.shape[1] == Na
            wa_p, wb_p = wa.data_ptr() + 4 * a0 * Ka, wb.data_ptr() + 4 * b0 * Na
            # W = W_b W_a  ([Nb][Na] x [Na][Ka]) and b = W_b b_a + b_b with the library's own fp32 GEMM
            self._gemm(Nb, Ka, Na, wb_p, Na, 1, wa_p, Ka, 1, mg["W"].data_ptr(), Ka)(st)
            bb_p = (mg["bb"].data_ptr() + 4 * b0) if mg["bb"] is not None else None
            if mg["ba"] is not None:
                self._gemm(Nb, 1, Na, wb_p, Na, 1, mg["ba"].data_ptr() + 4 * a0, 1, 1, mg["b"].data_ptr(), 1, res=bb_p, ldr=1)(st)
            elif mg["bb"] is not None:
                mg["b"].copy_(mg["bb"].detach()[b0:b1])
            else:
                mg["b"].zero_()
        self.flops = flops
        self._merged_version = ver

    def reseed(self, seed: int) -> None:
        """New dropout masks: every dropout site gets (seed, site id) as its Philox key / subsequence."""
        for s in self._drop_slots:
            s[0] = seed

    def load_inputs(self, x, t, cond) -> None:
        self.x_in.copy_(x)
        self.t_in.copy_(t)
        self.y_in.copy_(cond)

    def launch_forward(self) -> None:
        st = L.stream_ptr()
        if self._tapes is not None or self._ctapes is not None:      # eval mode at the reference batch sizes: ONE kernel (dense_cluster.cu / dense_fused.cu)
            self._launch_tape(st)
            return
        for _, fn in self.fwd_ops:
            fn(st)

    def launch(self) -> None:
        self.launch_forward()

    def launch_backward(self) -> None:
        st = L.stream_ptr()
        for _, fn in self.bwd_ops:
            fn(st)

    def num_launches(self):
        return len(self.fwd_ops), len(self.bwd_ops)

    def conv_flops(self) -> float:
        return self.flops


class _DenseTrainFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine: DenseEngine, x, t, y, *params):
        engine.load_inputs(x, t, y)
        engine.reseed(int(torch.randint(0, 2 ** 62, (1,)).item()) if engine._drop_slots else 0)
        engine.launch_forward()
        ctx.engine = engine
        engine.generation = getattr(engine, "generation", 0) + 1      # saved tensors are shared per batch size (train.py)
        ctx.generation = engine.generation
        return engine.eps.clone()

    @staticmethod
    def backward(ctx, d_eps):
        eng: DenseEngine = ctx.engine
        from .train import _check_generation
        _check_generation(eng, ctx)
        eng.d_eps.copy_(d_eps)
        eng.launch_backward()
        return (None, None, None, None) + tuple(eng.pgrad[k].clone() for k, _ in eng.module.named_parameters())


class DenseNoiseModel(CheckpointCompat, torch.nn.Module):
    """Shared dispatch of the two latent denoisers; subclasses declare parameters and ``_declare``."""
    in_dim = 20
    emb_mode = 0

    def _init_engines(self):
        self._engines: Dict[Tuple, DenseEngine] = {}
        self.precision = "fp32"

    def __getstate__(self):
        st = self.__dict__.copy()
        st["_engines"] = {}
        return st

    def _apply(self, fn, *a, **k):
        self._engines = {}
        return super()._apply(fn, *a, **k)

    def _declare(self, e: DenseEngine) -> None:
        raise NotImplementedError

    def engine(self, batch: int, device: torch.device, training: Optional[bool] = None) -> DenseEngine:
        training = self.training if training is None else training
        key = ("dense", batch, str(device), bool(training))
        eng = self._engines.get(key)
        if eng is None:
            eng = DenseEngine(self, batch, device, training, self.in_dim, self.emb_mode)
            self._declare(eng)
            eng.build()
            self._engines[key] = eng
        return eng

    def forward(self, x, t, y):
        device = L.require_device(x.device)
        p = next(self.parameters())
        if p.device != device:
            raise RuntimeError(f"NoiseModel parameters are on {p.device}, input on {device}")
        eng = self.engine(x.shape[0], device)
        x = x.to(torch.float32).contiguous()
        if self.training:
            params = tuple(q for _, q in self.named_parameters())
            return _DenseTrainFunction.apply(eng, x, t, y.to(device), *params)
        eng.load_inputs(x, t, y.to(device))
        eng.use_t_dev = False
        eng.refresh_weights()
        eng.launch_forward()
        return eng.eps.clone()


def dense_sample(vae, noise_model: DenseNoiseModel, diffusion, device, n_samples, y, x_T=None, z=None, seed=None,
                 use_graph=True, decode=True):
    """latent_diffusion.py:308-347 / diffusion_transformer.py:291-330."""
    from .process import ReverseLoop
    if y is None:
        raise ValueError("Class labels 'y' must be provided for conditional generation.")
    if y.shape[0] != n_samples:
        raise ValueError("y must have shape (n_samples,)")
    device = L.require_device(device)
    if vae is not None:
        vae.eval()
    noise_model.eval()
    latent = vae.config.latent_dim if vae is not None else noise_model.in_dim
    if x_T is None:
        x_T = torch.randn(n_samples, latent)                       # CPU generator, then H2D (:326)
    eng = noise_model.engine(n_samples, device, training=False)
    eng.x_in.copy_(x_T.to(torch.float32))
    eng.y_in.copy_(y)
    eng.use_t_dev = True
    eng.refresh_weights()
    loop = getattr(eng, "_reverse_loop", None)
    if loop is None or loop.p is not diffusion or loop.use_graph != use_graph:
        loop = ReverseLoop(diffusion, eng.x_in, eng.eps, eng.t_dev, eng.launch_forward, use_graph=use_graph,
                           fused_step=eng.fused_step if eng.can_fuse_step() else None)
        eng._reverse_loop = loop
    if z is not None:
        z = z.to(device=device, dtype=torch.float32).contiguous()
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    loop.run(z=z, seed=seed)
    zf = eng.x_in.clone()
    eng.use_t_dev = False
    if vae is None or not decode:
        return zf
    return vae.decode(zf).view(-1, 1, 28, 28)                       # :346
