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
                                     t_dev.data_ptr(), x.numel(), L.ptr(sd), L.stream_ptr()), "td_psample_step")
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
            y_nchw=False, out=None, y_coff=0, x_coff=0, cin=None):
    """x: NHWC (or NCHW fp32 when x_nchw).  Returns NHWC (or NCHW) output."""
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
    h = C.c_void_p()
    L.check(lib.td_conv3x3_plan_create(C.byref(h), C.byref(d), engine), "td_conv3x3_plan_create")
    try:
        L.check(lib.td_conv3x3_run(h, L.stream_ptr()), "td_conv3x3_run")
    finally:
        lib.td_conv3x3_plan_destroy(h)
    return out


def embed_head(t, w0, b0, w2, b2, proj_w, proj_b, in_mode=0, y=None, class_table=None, text=None):
    _dev(w0)
    B, D = t.shape[0], w2.shape[0]
    out = torch.empty(B, proj_w.shape[0], device=w0.device)
    emb = torch.empty(B, D, device=w0.device)
    a = L.EmbedArgs()
    a.batch, a.dim, a.in_mode, a.proj_out = B, D, in_mode, proj_w.shape[0]
    a.t, a.t_dev = t.data_ptr(), None
    a.w0, a.b0, a.w2, a.b2 = w0.data_ptr(), b0.data_ptr(), w2.data_ptr(), b2.data_ptr()
    a.y, a.class_table, a.text = L.ptr(y), L.ptr(class_table), L.ptr(text)
    a.proj_w, a.proj_b = proj_w.data_ptr(), proj_b.data_ptr()
    a.emb_out, a.h_out, a.proj_out_ptr = emb.data_ptr(), None, out.data_ptr()
    L.check(L.load().td_embed_head_fwd(C.byref(a), L.stream_ptr()), "td_embed_head_fwd")
    return out, emb
