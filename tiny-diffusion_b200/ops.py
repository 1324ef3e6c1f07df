"""Functional wrappers: one Python function per C-ABI entry point, taking/returning CUDA tensors.
These are what the kernel-level parity tests call; the model plans in ``unet.py`` call the C ABI
directly with cached pointers."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L


def _dev(t: torch.Tensor) -> torch.device:
    return L.require_device(t.device)


def qsample(x0, t, noise, alphas_cumprod, num_timesteps=None):
    _dev(x0)
    x0 = x0.contiguous().float()
    noise = noise.contiguous().float()
    x_t = torch.empty_like(x0)
    B = x0.shape[0]
    T = alphas_cumprod.numel() if num_timesteps is None else num_timesteps
    L.check(L.load().td_qsample(x0.data_ptr(), noise.data_ptr(), t.contiguous().data_ptr(),
                                alphas_cumprod.contiguous().data_ptr(), x_t.data_ptr(), B, x0.numel() // max(B, 1),
                                T, None, L.stream_ptr()), "td_qsample")
    return x_t


def qsample_philox(x0, t, alphas_cumprod, seed: int, offset: int = 0):
    _dev(x0)
    x0 = x0.contiguous().float()
    noise = torch.empty_like(x0)
    x_t = torch.empty_like(x0)
    sd = torch.tensor([seed, offset], dtype=torch.int64, device=x0.device)
    B = x0.shape[0]
    L.check(L.load().td_qsample(x0.data_ptr(), noise.data_ptr(), t.contiguous().data_ptr(),
                                alphas_cumprod.contiguous().data_ptr(), x_t.data_ptr(), B, x0.numel() // max(B, 1),
                                alphas_cumprod.numel(), sd.data_ptr(), L.stream_ptr()), "td_qsample")
    return x_t, noise


def mse_grad(pred, target, want_grad=True):
    _dev(pred)
    pred, target = pred.contiguous().float(), target.contiguous().float()
    n = pred.numel()
    lib = L.load()
    grad = torch.empty_like(pred) if want_grad else None
    loss = torch.zeros(1, device=pred.device)
    partials = torch.zeros(int(lib.td_mse_num_partials(n)), device=pred.device)
    counter = torch.zeros(1, device=pred.device, dtype=torch.int32)
    L.check(lib.td_mse_grad(pred.data_ptr(), target.data_ptr(), L.ptr(grad), loss.data_ptr(), partials.data_ptr(),
                            counter.data_ptr(), n, 1.0 / n, L.stream_ptr()), "td_mse_grad")
    return loss, grad


def psample_step(x, eps, z, coef, t: int, seed: Optional[int] = None):
    """In-place on x."""
    _dev(x)
    t_dev = torch.tensor([t], dtype=torch.int32, device=x.device)
    sd = None if seed is None else torch.tensor([seed, 0], dtype=torch.int64, device=x.device)
    L.check(L.load().td_psample_step(x.data_ptr(), eps.contiguous().data_ptr(), L.ptr(z), 0, coef.data_ptr(),
                                     t_dev.data_ptr(), x.numel(), coef.shape[0], L.ptr(sd), L.stream_ptr()), "td_psample_step")
    return x


def maxpool2(x, ceil_mode: bool):
    _dev(x)
    B, H, W, Cc = x.shape
    ho = (H + 1) // 2 if ceil_mode else H // 2
    wo = (W + 1) // 2 if ceil_mode else W // 2
    y = torch.empty(B, ho, wo, Cc, device=x.device, dtype=x.dtype)
    L.check(L.load().td_maxpool2_fwd(x.data_ptr(), y.data_ptr(), L.dtype_code(x.dtype), B, H, W, Cc, int(ceil_mode),
                                     L.stream_ptr()), "td_maxpool2_fwd")
    return y


def upcat(low, skip, temb, temb_off: int):
    _dev(low)
    B, hl, wl, cu = low.shape
    _, hs, ws, cs = skip.shape
    ho, wo = 2 * hl, 2 * wl
    out = torch.empty(B, ho, wo, cu + cs, device=low.device, dtype=low.dtype)
    L.check(L.load().td_upcat_fwd(low.data_ptr(), skip.data_ptr(), temb.data_ptr(), temb.shape[1], temb_off,
                                  out.data_ptr(), L.dtype_code(low.dtype), B, ho, wo, cu, hs, ws, cs, L.stream_ptr()),
            "td_upcat_fwd")
    return out


def resize_bilinear(x, ho: int, wo: int):
    _dev(x)
    B, hi, wi, Cc = x.shape
    y = torch.empty(B, ho, wo, Cc, device=x.device, dtype=x.dtype)
    L.check(L.load().td_resize_bilinear_fwd(x.data_ptr(), y.data_ptr(), L.dtype_code(x.dtype), B, hi, wi, ho, wo, Cc,
                                            L.stream_ptr()), "td_resize_bilinear_fwd")
    return y


def pack_conv_weight(w_oihw, dtype=torch.float32):
    _dev(w_oihw)
    co, ci = w_oihw.shape[:2]
    out = torch.empty(co, 3, 3, ci, device=w_oihw.device, dtype=dtype)
    L.check(L.load().td_pack_conv_weight(w_oihw.contiguous().data_ptr(), out.data_ptr(), L.dtype_code(dtype), co, ci,
                                         L.stream_ptr()), "td_pack_conv_weight")
    return out


def conv3x3(x, w_ohwi, scale=None, shift=None, relu=False, engine=L.CONV_SIMT, out_dtype=None, x_nchw=False,
            y_nchw=False, out=None, y_coff=0, x_coff=0, cin=None, want_stats=False, pool=None):
    """x: NHWC (or NCHW fp32 when x_nchw).  Returns NHWC (or NCHW) output.  ``pool`` = "ceil" / "floor": also returns
    ``(MaxPool2d(2) of the output or None, fused)`` -- the pooled map comes from the conv epilogue when the plan accepts it
    (``td_conv3x3_pool_fused``), else it is None and the caller pools the output itself."""
    _dev(x)
    lib = L.load()
    if x_nchw:
        B, ci_total, H, W = x.shape
    else:
        B, H, W, ci_total = x.shape
    cin = ci_total - x_coff if cin is None else cin
    cout = w_ohwi.shape[0]
    out_dtype = out_dtype or x.dtype
    if out is None:
        out = torch.empty((B, cout, H, W) if y_nchw else (B, H, W, cout), device=x.device, dtype=out_dtype)
    d = L.ConvDesc()
    d.batch, d.height, d.width, d.cin, d.cout = B, H, W, cin, cout
    d.x_dtype, d.y_dtype = L.dtype_code(x.dtype), L.dtype_code(out.dtype)
    d.x, d.ldx, d.x_coff = x.data_ptr(), (cin if x_nchw else ci_total), x_coff
    d.y, d.ldy, d.y_coff = out.data_ptr(), (cout if y_nchw else out.shape[3]), y_coff
    d.w, d.scale, d.shift = w_ohwi.data_ptr(), L.ptr(scale), L.ptr(shift)
    d.relu, d.stats, d.x_nchw, d.y_nchw = int(relu), None, int(x_nchw), int(y_nchw)
    ws = None
    if engine == L.CONV_TC:
        need = int(lib.td_conv3x3_splitk_workspace(C.byref(d)))
        if need > 0:
            ws = torch.empty(need, device=x.device)
            d.splitk_ws = ws.data_ptr()
    part = None
    if want_stats:
        part = torch.full(((B * H * W // 32 + 2) * 2 * cout + cout,), float("nan"), device=x.device)
        d.stats = part.data_ptr()
    pooled = None
    if pool is not None:
        ceil = pool == "ceil"
        Hp, Wp = ((H + 1) // 2, (W + 1) // 2) if ceil else (H // 2, W // 2)
        pooled = torch.full((B, Hp, Wp, cout), float("nan"), device=x.device, dtype=out.dtype)
        d.pool_y, d.pool_ceil = pooled.data_ptr(), int(ceil)
    h = C.c_void_p()
    L.check(lib.td_conv3x3_plan_create(C.byref(h), C.byref(d), engine), "td_conv3x3_plan_create")
    try:
        L.check(lib.td_conv3x3_run(h, L.stream_ptr()), "td_conv3x3_run")
        rows = int(lib.td_conv3x3_stats_rows(h)) if want_stats else 0
        fused = bool(lib.td_conv3x3_pool_fused(h)) if pool is not None else False
    finally:
        lib.td_conv3x3_plan_destroy(h)
    if pool is not None:
        return out, (pooled if fused else None)
    if want_stats:
        return out, part[:rows * 2 * cout].view(rows, 2, cout), part[rows * 2 * cout:rows * 2 * cout + cout]
    return out


def embed_head(t, w0, b0, w2, b2, proj_w, proj_b, in_mode=0, y=None, class_table=None, text=None, grads_for=None):
    """Returns (proj [B, P], emb [B, D]).  With ``grads_for=d_proj`` also returns the parameter gradients."""
    _dev(w0)
    lib = L.load()
    B, D = t.shape[0], w2.shape[0]
    out = torch.empty(B, proj_w.shape[0], device=w0.device)
    saved = torch.empty(int(lib.td_embed_head_saved_floats(B, D, in_mode)), device=w0.device)
    a = L.EmbedArgs()
    a.batch, a.dim, a.in_mode, a.proj_out = B, D, in_mode, proj_w.shape[0]
    a.t, a.t_dev = t.data_ptr(), None
    a.w0, a.b0, a.w2, a.b2 = w0.data_ptr(), b0.data_ptr(), w2.data_ptr(), b2.data_ptr()
    a.y, a.class_table, a.text = L.ptr(y), L.ptr(class_table), L.ptr(text)
    a.proj_w, a.proj_b = proj_w.data_ptr(), proj_b.data_ptr()
    a.saved, a.proj_out_ptr = saved.data_ptr(), out.data_ptr()
    L.check(lib.td_embed_head_fwd(C.byref(a), L.stream_ptr()), "td_embed_head_fwd")
    din = D if in_mode == 2 else 1
    emb = saved[B * (din + 2 * D):].view(B, D).clone()
    if grads_for is None:
        return out, emb
    g = L.EmbedGrads()
    d_proj = grads_for.contiguous()
    scratch = torch.empty(2 * B * D, device=w0.device)
    res = {"w0": torch.empty_like(w0), "b0": torch.empty_like(b0), "w2": torch.empty_like(w2),
           "b2": torch.empty_like(b2), "proj_w": torch.empty_like(proj_w), "proj_b": torch.empty_like(proj_b)}
    if class_table is not None:
        res["class_table"] = torch.empty_like(class_table)
    g.d_proj, g.scratch = d_proj.data_ptr(), scratch.data_ptr()
    g.d_w0, g.d_b0, g.d_w2, g.d_b2 = (res[k].data_ptr() for k in ("w0", "b0", "w2", "b2"))
    g.d_class_table = res["class_table"].data_ptr() if class_table is not None else None
    g.num_classes = class_table.shape[0] if class_table is not None else 0
    g.d_proj_w, g.d_proj_b = res["proj_w"].data_ptr(), res["proj_b"].data_ptr()
    L.check(lib.td_embed_head_bwd(C.byref(a), C.byref(g), L.stream_ptr()), "td_embed_head_bwd")
    return out, emb, res


# ---------------------------------------------------------------------------------------------
# training kernels
# ---------------------------------------------------------------------------------------------
def pack_conv_weight_dgrad(w_oihw, dtype=torch.float32):
    _dev(w_oihw)
    co, ci = w_oihw.shape[:2]
    out = torch.empty(ci, 3, 3, co, device=w_oihw.device, dtype=dtype)
    L.check(L.load().td_pack_conv_weight_dgrad(w_oihw.contiguous().data_ptr(), out.data_ptr(), L.dtype_code(dtype),
                                               co, ci, L.stream_ptr()), "td_pack_conv_weight_dgrad")
    return out


def final_resize_conv(x, w_ohwi, bias, ho, wo):
    """x: NHWC (bf16/fp32); w_ohwi: fp32 [1,3,3,C]; returns fp32 NCHW [B,1,ho,wo] = final_conv(interpolate(x, (ho,wo)))."""
    _dev(x)
    B, hi, wi, c = x.shape
    out = torch.empty(B, 1, ho, wo, device=x.device, dtype=torch.float32)
    L.check(L.load().td_final_resize_conv(x.data_ptr(), L.dtype_code(x.dtype), c, 0, B, hi, wi, c, w_ohwi.data_ptr(),
                                          L.ptr(bias), ho, wo, out.data_ptr(), L.stream_ptr()), "td_final_resize_conv")
    return out


def conv3x3_wgrad(x, dy, engine=L.CONV_SIMT, x_nchw=False, dy_nchw=False, x_coff=0, cin=None):
    """x: conv input NHWC (or NCHW fp32), dy: grad of the conv output NHWC (or NCHW).  Returns OIHW fp32."""
    _dev(x)
    lib = L.load()
    if x_nchw:
        B, ci_total, H, W = x.shape
    else:
        B, H, W, ci_total = x.shape
    cin = ci_total - x_coff if cin is None else cin
    cout = dy.shape[1] if dy_nchw else dy.shape[3]
    dw = torch.empty(cout, cin, 3, 3, device=x.device, dtype=torch.float32)
    d = L.WgradDesc()
    d.batch, d.height, d.width, d.cin, d.cout = B, H, W, cin, cout
    d.x_dtype, d.dy_dtype = L.dtype_code(x.dtype), L.dtype_code(dy.dtype)
    d.x, d.ldx, d.x_coff, d.x_nchw = x.data_ptr(), (cin if x_nchw else ci_total), x_coff, int(x_nchw)
    d.dy, d.lddy, d.dy_coff, d.dy_nchw = dy.data_ptr(), cout, 0, int(dy_nchw)
    d.dw = dw.data_ptr()
    ws = torch.empty(max(int(lib.td_conv3x3_wgrad_workspace(C.byref(d), engine)), 1), device=x.device)
    d.workspace = ws.data_ptr()
    h = C.c_void_p()
    L.check(lib.td_conv3x3_wgrad_plan_create(C.byref(h), C.byref(d), engine), "td_conv3x3_wgrad_plan_create")
    try:
        L.check(lib.td_conv3x3_wgrad_run(h, L.stream_ptr()), "td_conv3x3_wgrad_run")
    finally:
        lib.td_conv3x3_wgrad_plan_destroy(h)
    return dw


def bn_train_fwd(y, gamma, beta, conv_bias, running_mean, running_var, nbt, eps=1e-5, momentum=0.1, relu=True,
                 out_dtype=torch.float32, fused=False):
    """y: raw fp32 conv output NHWC (bias excluded).  Returns (a, scale, shift, save_mean, save_invstd).
    ``fused``: td_bn_apply_fused (finalize in the prologue of the apply pass) instead of finalize + apply."""
    _dev(y)
    assert y.dtype == torch.float32
    lib = L.load()
    Cc = y.shape[-1]
    P = y.numel() // Cc
    rows = int(lib.td_chan_reduce_rows(L.TD_F32, P, Cc))
    part = torch.empty(rows * 2 * Cc + Cc, device=y.device)
    st = L.stream_ptr()
    L.check(lib.td_bn_stats(y.data_ptr(), L.TD_F32, Cc, 0, P, Cc, part.data_ptr(), 1, st), "td_bn_stats")
    scale, shift, mean, invstd = (torch.empty(Cc, device=y.device) for _ in range(4))
    if fused:
        a = torch.empty(y.shape, device=y.device, dtype=out_dtype)
        L.check(lib.td_bn_apply_fused(y.data_ptr(), L.TD_F32, part.data_ptr(), rows, P, gamma.data_ptr(), beta.data_ptr(),
                                      L.ptr(conv_bias), eps, momentum, L.ptr(running_mean), L.ptr(running_var), L.ptr(nbt),
                                      scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), invstd.data_ptr(), a.data_ptr(),
                                      L.dtype_code(out_dtype), Cc, 0, P, Cc, int(relu), st), "td_bn_apply_fused")
        return a, scale, shift, mean, invstd
    L.check(lib.td_bn_finalize(part.data_ptr(), rows, Cc, P, gamma.data_ptr(), beta.data_ptr(), L.ptr(conv_bias), eps,
                               momentum, L.ptr(running_mean), L.ptr(running_var), L.ptr(nbt), scale.data_ptr(),
                               shift.data_ptr(), mean.data_ptr(), invstd.data_ptr(), st), "td_bn_finalize")
    a = torch.empty(y.shape, device=y.device, dtype=out_dtype)
    L.check(lib.td_bn_relu_apply(y.data_ptr(), scale.data_ptr(), shift.data_ptr(), a.data_ptr(), L.dtype_code(out_dtype),
                                 Cc, 0, P, Cc, int(relu), st), "td_bn_relu_apply")
    return a, scale, shift, mean, invstd


def bn_train_bwd(da, y, scale, shift, mean, invstd, fused=False, beta=None):
    """da: grad of the block output (fp32 / bf16), y: raw fp32 conv output.  Returns (dy, dgamma, dbeta).
    ``fused``: td_bn_bwd_reduce + td_bn_bwd_apply_fused instead of reduce + finalize + apply."""
    _dev(y)
    lib = L.load()
    Cc = y.shape[-1]
    P = y.numel() // Cc
    dt = L.dtype_code(da.dtype)
    if fused:
        rows = int(lib.td_bn_bwd_reduce_rows(dt, P, Cc))
        part = torch.empty(rows * 2 * Cc + Cc, device=y.device)
        st = L.stream_ptr()
        L.check(lib.td_bn_bwd_reduce(da.data_ptr(), Cc, 0, y.data_ptr(), L.dtype_code(y.dtype), dt, scale.data_ptr(), beta.data_ptr(),
                                     mean.data_ptr(), P, Cc, part.data_ptr(), st), "td_bn_bwd_reduce")
        dgamma, dbeta = torch.empty(Cc, device=y.device), torch.empty(Cc, device=y.device)
        dy = torch.empty_like(da)
        L.check(lib.td_bn_bwd_apply_fused(da.data_ptr(), Cc, 0, y.data_ptr(), L.dtype_code(y.dtype), dt, part.data_ptr(), rows, P, scale.data_ptr(),
                                          beta.data_ptr(), mean.data_ptr(), invstd.data_ptr(), dgamma.data_ptr(),
                                          dbeta.data_ptr(), dy.data_ptr(), P, Cc, st), "td_bn_bwd_apply_fused")
        return dy, dgamma, dbeta
    rows = int(lib.td_chan_reduce_rows(dt, P, Cc))
    part = torch.empty(rows * 2 * Cc + Cc, device=y.device)
    st = L.stream_ptr()
    L.check(lib.td_bn_relu_bwd_reduce(da.data_ptr(), Cc, 0, y.data_ptr(), dt, scale.data_ptr(), shift.data_ptr(),
                                      mean.data_ptr(), P, Cc, part.data_ptr(), st), "td_bn_relu_bwd_reduce")
    dgamma, dbeta = torch.empty(Cc, device=y.device), torch.empty(Cc, device=y.device)
    coef = torch.empty(3, Cc, device=y.device)
    L.check(lib.td_bn_bwd_finalize(part.data_ptr(), rows, Cc, P, scale.data_ptr(), mean.data_ptr(), invstd.data_ptr(),
                                   dgamma.data_ptr(), dbeta.data_ptr(), coef.data_ptr(), st), "td_bn_bwd_finalize")
    dy = torch.empty_like(da)
    L.check(lib.td_bn_relu_bwd_apply(da.data_ptr(), Cc, 0, y.data_ptr(), dt, scale.data_ptr(), shift.data_ptr(),
                                     coef.data_ptr(), dy.data_ptr(), P, Cc, st), "td_bn_relu_bwd_apply")
    return dy, dgamma, dbeta


def maxpool2_bwd(x, dy, ceil_mode: bool, dx=None):
    _dev(x)
    B, H, W, Cc = x.shape
    acc = dx is not None
    if dx is None:
        dx = torch.empty_like(x)
    L.check(L.load().td_maxpool2_bwd(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), L.dtype_code(x.dtype), B, H, W, Cc,
                                     int(ceil_mode), int(acc), L.stream_ptr()), "td_maxpool2_bwd")
    return dx


def resize_bilinear_bwd(dy, hi: int, wi: int):
    _dev(dy)
    B, ho, wo, Cc = dy.shape
    dx = torch.empty(B, hi, wi, Cc, device=dy.device, dtype=dy.dtype)
    L.check(L.load().td_resize_bilinear_bwd(dy.data_ptr(), Cc, 0, dx.data_ptr(), L.dtype_code(dy.dtype), B, hi, wi, ho,
                                            wo, Cc, L.stream_ptr()), "td_resize_bilinear_bwd")
    return dx


def upcat_bwd(dout, cu: int, hs: int, ws: int, ld_temb: int, temb_off: int):
    """Returns (dlow, dskip, dtemb [B, ld_temb] with only [temb_off, temb_off+cs) written)."""
    _dev(dout)
    B, ho, wo, ct = dout.shape
    cs = ct - cu
    dlow = torch.empty(B, ho // 2, wo // 2, cu, device=dout.device, dtype=dout.dtype)
    dskip = torch.empty(B, hs, ws, cs, device=dout.device, dtype=dout.dtype)
    dtemb = torch.zeros(B, ld_temb, device=dout.device)
    L.check(L.load().td_upcat_bwd(dout.data_ptr(), dlow.data_ptr(), dskip.data_ptr(), dtemb.data_ptr(), ld_temb,
                                  temb_off, L.dtype_code(dout.dtype), B, ho, wo, cu, hs, ws, cs, L.stream_ptr()),
            "td_upcat_bwd")
    return dlow, dskip, dtemb


def gemm(A, B, bias=None, act=L.ACT_NONE, residual=None, trans_a=False, trans_b=False, splitk=False):
    """C = act(op(A) @ op(B) + bias) + residual (fp32).  A: [M,K] (or [K,M] with trans_a), B: [K,N] (or [N,K])."""
    _dev(A)
    lib = L.load()
    M, K = (A.shape[1], A.shape[0]) if trans_a else A.shape
    N = B.shape[0] if trans_b else B.shape[1]
    out = torch.empty(M, N, device=A.device)
    g = L.GemmArgs()
    g.M, g.N, g.K, g.alpha = M, N, K, 1.0
    g.A, g.a_rs, g.a_cs = A.data_ptr(), (1 if trans_a else A.stride(0)), (A.stride(0) if trans_a else 1)
    g.B, g.b_rs, g.b_cs = B.data_ptr(), (1 if trans_b else B.stride(0)), (B.stride(0) if trans_b else 1)
    g.C, g.ldc, g.bias, g.act = out.data_ptr(), N, L.ptr(bias), act
    g.residual, g.ldr = L.ptr(residual), N
    ws = None
    if splitk:
        n = int(lib.td_gemm_f32_workspace(M, N, K))
        if n > 0:
            ws = torch.empty(n, device=A.device)
            g.splitk_ws = ws.data_ptr()
    L.check(lib.td_gemm_f32(C.byref(g), L.stream_ptr()), "td_gemm_f32")
    return out
