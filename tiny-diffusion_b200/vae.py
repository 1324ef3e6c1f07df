"""The MLP VAE at the edges of the latent pipelines (vae.py:16-67): ``encode`` before ``q_sample``
(latent_diffusion.py:207-209) and ``decode`` after the reverse loop (:346), on libtinydiff GEMMs.
Inference-only (the diffusion scripts call these under ``torch.no_grad``); training the VAE itself is
out of scope."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch
import torch.nn as nn

from . import _lib as L


@dataclass
class VAEConfig:
    """vae.py:16-25 (the fields the diffusion scripts read)."""
    input_dim: int = 784
    hidden_dim: int = 400
    latent_dim: int = 20


def _linear(x: torch.Tensor, lin: nn.Linear, act: int) -> torch.Tensor:
    x = x.contiguous()
    M, K = x.shape
    N = lin.weight.shape[0]
    out = torch.empty(M, N, device=x.device)
    g = L.GemmArgs()
    g.M, g.N, g.K, g.alpha = M, N, K, 1.0
    g.A, g.a_rs, g.a_cs = x.data_ptr(), K, 1
    g.B, g.b_rs, g.b_cs = lin.weight.data_ptr(), 1, K
    g.C, g.ldc, g.bias, g.act = out.data_ptr(), N, lin.bias.data_ptr(), act
    L.check(L.load().td_gemm_f32(C.byref(g), L.stream_ptr()), "td_gemm_f32")
    return out


class VAE(nn.Module):
    def __init__(self, config: VAEConfig = None):
        super().__init__()
        self.config = config or VAEConfig()
        c = self.config
        self.fc1 = nn.Linear(c.input_dim, c.hidden_dim)
        self.fc21 = nn.Linear(c.hidden_dim, c.latent_dim)
        self.fc22 = nn.Linear(c.hidden_dim, c.latent_dim)
        self.fc3 = nn.Linear(c.latent_dim, c.hidden_dim)
        self.fc4 = nn.Linear(c.hidden_dim, c.input_dim)

    @torch.no_grad()
    def encode(self, x):
        L.require_device(x.device)
        h1 = _linear(x.float(), self.fc1, L.ACT_RELU)
        return _linear(h1, self.fc21, L.ACT_NONE), _linear(h1, self.fc22, L.ACT_NONE)

    @torch.no_grad()
    def reparameterize(self, mu, logvar, eps=None):
        """vae.py:55-58; ``eps`` may be injected."""
        std = torch.exp(0.5 * logvar)
        eps = torch.randn_like(std) if eps is None else eps
        return mu + eps * std

    @torch.no_grad()
    def decode(self, z):
        L.require_device(z.device)
        return _linear(_linear(z.float(), self.fc3, L.ACT_RELU), self.fc4, L.ACT_SIGMOID)

    def forward(self, x):
        mu, logvar = self.encode(x.view(-1, self.config.input_dim))
        return self.decode(self.reparameterize(mu, logvar)), mu, logvar
