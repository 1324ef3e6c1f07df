"""Drop-in for the encode / decode path of the reference's conv / attention VAE, ``vae_laion.py``
(``VAE`` :88-203, ``SelfAttention`` :50-65, ``ResidualBlock`` :69-85) -- SURVEY.md 8f #3: the only home of
``ConvTranspose2d`` and spatial self-attention in the reference.  Same constructor, ``state_dict`` layout
(``weight_orig`` / ``weight_u`` / ``weight_v`` of ``torch.nn.utils.spectral_norm``) and default initialisation; the
``torch.nn`` sub-modules only own the parameters and are never called.

    from tinydiff.vae_laion import VAE, VAEConfig
    mu, logvar = vae.encode(x)            # x: (B, 3, 256, 256) in [0, 1]
    recon = vae.decode(z)                 # (B, 3, 256, 256)

Inference (eval mode) only: the VAE's training objective needs pretrained VGG16 features (:171-176, not available offline)
and is not a diffusion hot path.  Kernels (libtinydiff, fp32, NHWC): stride-2 4x4 convolution and its transpose as
gather-form implicit GEMMs, 3x3 residual convolutions with the BatchNorm folded into the epilogue, flash-style attention
(the (HW)^2 matrix -- 1 GiB per image at 128 x 128 -- is never materialised), spectral-norm sigma (with power iteration).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Any, Dict, List, Optional

import torch
import torch.nn as nn

from . import _lib as L

__all__ = ["VAE", "VAEConfig", "SelfAttention", "ResidualBlock"]


@dataclass
class VAEConfig:
    """vae_laion.py:25-39 (the fields the model reads; the data-loading / logging fields are kept for signature parity)."""
    latent_dim: int = 128
    hidden_channels: int = 64
    input_channels: int = 3
    image_size: int = 256
    batch_size: int = 4
    epochs: int = 100
    learning_rate: float = 1e-4
    device: Any = None
    checkpoint_dir: str = "checkpoints"
    beta: float = 1.0


class SelfAttention(nn.Module):
    """Parameter layout of vae_laion.py:51-55."""

    def __init__(self, in_channels: int):
        super().__init__()
        self.query = nn.Conv2d(in_channels, in_channels // 8, 1)
        self.key = nn.Conv2d(in_channels, in_channels // 8, 1)
        self.value = nn.Conv2d(in_channels, in_channels, 1)
        self.gamma = nn.Parameter(torch.zeros(1))


class ResidualBlock(nn.Module):
    """Parameter layout of vae_laion.py:70-78."""

    def __init__(self, channels: int):
        super().__init__()
        self.conv1 = nn.utils.spectral_norm(nn.Conv2d(channels, channels, 3, padding=1, bias=False))
        self.bn1 = nn.BatchNorm2d(channels)
        self.conv2 = nn.utils.spectral_norm(nn.Conv2d(channels, channels, 3, padding=1, bias=False))
        self.bn2 = nn.BatchNorm2d(channels)


def _gemm(M, N, K, A, a_rs, Bm, b_rs, b_cs, Cm, ldc, bias=None, act=L.ACT_NONE, ws: Optional[torch.Tensor] = None):
    g = L.GemmArgs()
    g.M, g.N, g.K, g.alpha = M, N, K, 1.0
    g.A, g.a_rs, g.a_cs = A, a_rs, 1
    g.B, g.b_rs, g.b_cs = Bm, b_rs, b_cs
    g.C, g.ldc, g.bias, g.act = Cm, ldc, bias, act
    g.splitk_ws = L.ptr(ws)
    L.check(L.load().td_gemm_f32(C.byref(g), L.stream_ptr()), "td_gemm_f32")


class VAE(nn.Module):
    ENC = [(3, 32, True), (32, 64, True), (64, 128, False), (128, 256, False)]       # cin, cout, attention (vae_laion.py:95-133)
    DEC = [(256, 128, True), (128, 64, True), (64, 32, False)]                       # :136-160 (+ the output layer :161-167)

    def __init__(self, config: Optional[VAEConfig] = None):
        super().__init__()
        self.config = config or VAEConfig()
        c = self.config
        sn = nn.utils.spectral_norm
        enc = []
        for cin, cout, attn in self.ENC:
            cin = c.input_channels if cin == 3 else cin
            mods = [sn(nn.Conv2d(cin, cout, 4, stride=2, padding=1)), nn.ReLU(), ResidualBlock(cout)]
            if attn:
                mods.append(SelfAttention(cout))
            enc.append(nn.Sequential(*mods))
        self.encoder = nn.ModuleList(enc)
        self.fc_mu = nn.Linear(256 * 16 * 16, c.latent_dim)
        self.fc_logvar = nn.Linear(256 * 16 * 16, c.latent_dim)
        self.decoder_input = nn.Linear(c.latent_dim, 256 * 16 * 16)
        dec = []
        for cin, cout, attn in self.DEC:
            mods = [sn(nn.ConvTranspose2d(cin, cout, 4, stride=2, padding=1)), nn.ReLU(), ResidualBlock(cout)]
            if attn:
                mods.append(SelfAttention(cout))
            dec.append(nn.Sequential(*mods))
        dec.append(nn.Sequential(sn(nn.ConvTranspose2d(32, c.input_channels, 4, stride=2, padding=1)), nn.Sigmoid()))
        self.decoder = nn.ModuleList(dec)
        self._packed: Dict[str, torch.Tensor] = {}
        self._version = None

    def _apply(self, fn, *a, **k):
        self._packed, self._version = {}, None
        return super()._apply(fn, *a, **k)

    # ------------------------------------------------------------------ parameter preparation
    def _weights_version(self):
        return tuple(p._version for p in self.parameters()) + tuple(b._version for b in self.buffers()) + (
            getattr(self, "_weights_gen", 0),)

    def _sigma(self, mod: nn.Module, dim1: bool, power_iterations: int = 0) -> torch.Tensor:
        """sigma of one spectral-norm layer from its ``weight_orig`` / ``weight_u`` / ``weight_v`` (td_spectral_sigma)."""
        w = mod.weight_orig.detach()
        rows = w.shape[1] if dim1 else w.shape[0]
        cols = w.numel() // rows
        khw = w.shape[2] * w.shape[3]
        dev = w.device
        sig = torch.empty(1, device=dev)
        scratch = torch.empty(rows, device=dev)
        L.check(L.load().td_spectral_sigma(w.data_ptr(), rows, cols, int(dim1), khw, mod.weight_u.data_ptr(),
                                           mod.weight_v.data_ptr(), power_iterations, 1e-12, sig.data_ptr(), scratch.data_ptr(),
                                           L.stream_ptr()), "td_spectral_sigma")
        return sig

    def refresh_weights(self, force: bool = False) -> None:
        """Pack every operand once per weight version: 4x4 layers as [Cout][taps*Cin] / sigma, 3x3 layers OHWI with 1/sigma and
        the eval-mode BatchNorm folded into the epilogue affine, q|k|v as one [C/4 + C][C] matrix, the Linear layers
        re-ordered for NHWC flattening."""
        ver = self._weights_version()
        if not force and ver == self._version and self._packed:
            return
        lib, st = L.load(), L.stream_ptr()
        pk: Dict[str, torch.Tensor] = {}
        dev = self.fc_mu.weight.device

        def pack4(name, mod, transposed):
            w = mod.weight_orig.detach()
            cout, cin = (w.shape[1], w.shape[0]) if transposed else (w.shape[0], w.shape[1])
            sig = self._sigma(mod, transposed)
            out = torch.empty((4 if transposed else 1) * cout * (4 if transposed else 16) * cin, device=dev)
            L.check(lib.td_pack_conv4x4_weight(w.data_ptr(), sig.data_ptr(), out.data_ptr(), cout, cin, int(transposed), st),
                    "td_pack_conv4x4_weight")
            pk[name] = out

        def res(name, blk: ResidualBlock):
            for i, (conv, bn) in enumerate(((blk.conv1, blk.bn1), (blk.conv2, blk.bn2)), start=1):
                w = conv.weight_orig.detach()
                ch = w.shape[0]
                sig = self._sigma(conv, False)
                wp = torch.empty(ch, 3, 3, ch, device=dev)
                L.check(lib.td_pack_conv_weight(w.data_ptr(), wp.data_ptr(), L.TD_F32, ch, ch, st), "td_pack_conv_weight")
                scale, shift = torch.empty(ch, device=dev), torch.empty(ch, device=dev)
                L.check(lib.td_bn_fold(bn.weight.data_ptr(), bn.bias.data_ptr(), bn.running_mean.data_ptr(),
                                       bn.running_var.data_ptr(), None, float(bn.eps), scale.data_ptr(), shift.data_ptr(), ch, st),
                        "td_bn_fold")
                L.check(lib.td_scale_by_inv_sigma(scale.data_ptr(), sig.data_ptr(), ch, st), "td_scale_by_inv_sigma")
                pk[f"{name}.w{i}"], pk[f"{name}.scale{i}"], pk[f"{name}.shift{i}"] = wp, scale, shift

        def attn(name, a: SelfAttention):
            ch = a.value.weight.shape[0]
            pk[name + ".wqkv"] = torch.cat([a.query.weight.detach().view(-1, ch), a.key.weight.detach().view(-1, ch),
                                            a.value.weight.detach().view(-1, ch)]).contiguous()
            pk[name + ".bqkv"] = torch.cat([a.query.bias.detach(), a.key.bias.detach(), a.value.bias.detach()]).contiguous()

        for i, seq in enumerate(self.encoder):
            pack4(f"enc{i}", seq[0], False)
            res(f"enc{i}.res", seq[2])
            if len(seq) > 3:
                attn(f"enc{i}.attn", seq[3])
        for i, seq in enumerate(self.decoder):
            pack4(f"dec{i}", seq[0], True)
            if i < 3:
                res(f"dec{i}.res", seq[2])
                if len(seq) > 3:
                    attn(f"dec{i}.attn", seq[3])
        # h.view(B, -1) flattens NCHW (vae_laion.py:181,193); the activations here are NHWC: permute the Linear weights once
        lat = self.config.latent_dim
        for nm, lin in (("fc_mu", self.fc_mu), ("fc_logvar", self.fc_logvar)):
            pk[nm] = lin.weight.detach().view(lat, 256, 16, 16).permute(0, 2, 3, 1).reshape(lat, -1).contiguous()
        pk["fc_b"] = torch.cat([self.fc_mu.bias.detach(), self.fc_logvar.bias.detach()]).contiguous()
        pk["fc_w"] = torch.cat([pk.pop("fc_mu"), pk.pop("fc_logvar")]).contiguous()                     # [2*lat][65536]
        pk["dec_in_w"] = self.decoder_input.weight.detach().view(256, 16, 16, lat).permute(1, 2, 0, 3).reshape(-1, lat).contiguous()
        pk["dec_in_b"] = self.decoder_input.bias.detach().view(256, 16, 16).permute(1, 2, 0).reshape(-1).contiguous()
        self._packed, self._version = pk, ver

    # ------------------------------------------------------------------ building blocks (NHWC fp32)
    def _res_block(self, name: str, x: torch.Tensor) -> torch.Tensor:
        """vae_laion.py:80-85: relu(bn1(conv1 x)) -> bn2(conv2 .) + x, eval-mode BatchNorm folded into the conv epilogue."""
        from . import ops
        pk = self._packed
        y = ops.conv3x3(x, pk[f"{name}.w1"], pk[f"{name}.scale1"], pk[f"{name}.shift1"], relu=True, engine=L.CONV_SIMT)
        z = ops.conv3x3(y, pk[f"{name}.w2"], pk[f"{name}.scale2"], pk[f"{name}.shift2"], relu=False, engine=L.CONV_SIMT)
        Bn, H, W, Cc = x.shape
        L.check(L.load().td_add2d_f32(x.data_ptr(), Cc, z.data_ptr(), Cc, Bn * H * W, Cc, 1, L.stream_ptr()), "td_add2d_f32")
        return z

    def _attention(self, name: str, a: SelfAttention, x: torch.Tensor) -> torch.Tensor:
        pk = self._packed
        Bn, H, W, Cc = x.shape
        n, dq = H * W, Cc // 8
        ld = 2 * dq + Cc
        qkv = torch.empty(Bn * n, ld, device=x.device)
        w = pk[name + ".wqkv"]
        _gemm(Bn * n, ld, Cc, x.data_ptr(), Cc, w.data_ptr(), 1, Cc, qkv.data_ptr(), ld, bias=pk[name + ".bqkv"].data_ptr())
        y = torch.empty_like(x)
        L.check(L.load().td_self_attention_fwd(qkv.data_ptr(), x.data_ptr(), a.gamma.data_ptr(), y.data_ptr(), Bn, n, dq, Cc,
                                               L.stream_ptr()), "td_self_attention_fwd")
        return y

    def _check_mode(self):
        if self.training:
            raise NotImplementedError("tinydiff.vae_laion.VAE runs in eval mode only (call .eval()): training this VAE needs the "
                                      "pretrained VGG16 perceptual loss of vae_laion.py:171-176 and is out of scope")

    # ------------------------------------------------------------------ public API (vae_laion.py:177-203)
    @torch.no_grad()
    def encode(self, x: torch.Tensor):
        self._check_mode()
        dev = L.require_device(x.device)
        self.refresh_weights()
        lib, st, pk = L.load(), L.stream_ptr(), self._packed
        x = x.to(torch.float32).contiguous()
        Bn, cin, H, W = x.shape
        h = x
        nchw = 1
        for i, seq in enumerate(self.encoder):
            cout = seq[0].weight_orig.shape[0]
            y = torch.empty(Bn, H // 2, W // 2, cout, device=dev)
            L.check(lib.td_conv4x4s2_fwd(h.data_ptr(), pk[f"enc{i}"].data_ptr(), seq[0].bias.data_ptr(), y.data_ptr(), Bn, H, W, cin,
                                         cout, nchw, L.ACT_RELU, st), "td_conv4x4s2_fwd")
            h, cin, H, W, nchw = y, cout, H // 2, W // 2, 0
            h = self._res_block(f"enc{i}.res", h)
            if len(seq) > 3:
                h = self._attention(f"enc{i}.attn", seq[3], h)
        lat = self.config.latent_dim
        K = h.numel() // Bn
        out = torch.empty(Bn, 2 * lat, device=dev)
        need = int(lib.td_gemm_f32_workspace(Bn, 2 * lat, K))
        ws = torch.empty(need, device=dev) if need > 0 else None
        _gemm(Bn, 2 * lat, K, h.data_ptr(), K, pk["fc_w"].data_ptr(), 1, K, out.data_ptr(), 2 * lat, bias=pk["fc_b"].data_ptr(), ws=ws)
        return out[:, :lat].contiguous(), out[:, lat:].contiguous()

    @torch.no_grad()
    def reparameterize(self, mu, logvar, eps=None):
        """vae_laion.py:186-189; ``eps`` may be injected."""
        std = torch.exp(0.5 * logvar)
        eps = torch.randn_like(std) if eps is None else eps
        return mu + eps * std

    @torch.no_grad()
    def decode(self, z: torch.Tensor) -> torch.Tensor:
        self._check_mode()
        dev = L.require_device(z.device)
        self.refresh_weights()
        lib, st, pk = L.load(), L.stream_ptr(), self._packed
        z = z.to(torch.float32).contiguous()
        Bn, lat = z.shape
        h = torch.empty(Bn, 16, 16, 256, device=dev)
        _gemm(Bn, 256 * 16 * 16, lat, z.data_ptr(), lat, pk["dec_in_w"].data_ptr(), 1, lat, h.data_ptr(), 256 * 16 * 16,
              bias=pk["dec_in_b"].data_ptr())
        H = W = 16
        cin = 256
        for i, seq in enumerate(self.decoder):
            cout = seq[0].weight_orig.shape[1]
            last = i == 3
            y = torch.empty((Bn, cout, 2 * H, 2 * W) if last else (Bn, 2 * H, 2 * W, cout), device=dev)
            L.check(lib.td_convT4x4s2_fwd(h.data_ptr(), pk[f"dec{i}"].data_ptr(), seq[0].bias.data_ptr(), y.data_ptr(), Bn, H, W, cin,
                                          cout, int(last), L.ACT_SIGMOID if last else L.ACT_RELU, st), "td_convT4x4s2_fwd")
            h, cin, H, W = y, cout, 2 * H, 2 * W
            if last:
                return h
            h = self._res_block(f"dec{i}.res", h)
            if len(seq) > 3:
                h = self._attention(f"dec{i}.attn", seq[3], h)
        return h

    def forward(self, x):
        mu, logvar = self.encode(x)
        z = self.reparameterize(mu, logvar)
        return self.decode(z), mu, logvar
