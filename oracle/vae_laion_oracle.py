"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the conv / attention VAE of the reference's ``vae_laion.py``:
a functional restatement of ``VAE.encode`` / ``VAE.decode`` (vae_laion.py:88-203) -- spectral-norm stride-2 4x4 convolutions
(:95-133), ``ConvTranspose2d`` decoder (:138-168), ``ResidualBlock`` (:69-85), full HWxHW ``SelfAttention`` (:50-65) -- over a
``state_dict`` laid out exactly like the reference's (``weight_orig`` / ``weight_u`` / ``weight_v`` of
``torch.nn.utils.spectral_norm``).  Only tests/ import this module; the product path never does.

Pinned against the UNMODIFIED reference class in the build container (tests/test_oracle_vs_reference.py, through
oracle/reference_shim.py with the HF dataset, the VGG16 download and the hard-coded CUDA device stubbed) and against the
reference-generated golden outputs in tests/golden/vae_laion.pt on every box."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]


def spectral_weight(sd: StateDict, prefix: str, dim: int = 0, training: bool = False, eps: float = 1e-12,
                    new_uv: dict | None = None) -> Tensor:
    """torch.nn.utils.spectral_norm's ``compute_weight`` (one power iteration in training mode, none in eval):
    W / sigma with sigma = u^T W_mat v, W_mat = weight_orig with ``dim`` moved first and flattened (dim = 1 for
    ConvTranspose2d).  vae_laion.py:72-78,98-131,138-165."""
    w = sd[prefix + ".weight_orig"]
    u, v = sd[prefix + ".weight_u"], sd[prefix + ".weight_v"]
    wm = w
    if dim != 0:
        wm = wm.permute(dim, *[d for d in range(wm.dim()) if d != dim])
    wm = wm.reshape(wm.shape[0], -1)
    if training:
        with torch.no_grad():
            v = F.normalize(torch.mv(wm.t(), u), dim=0, eps=eps)
            u = F.normalize(torch.mv(wm, v), dim=0, eps=eps)
        if new_uv is not None:
            new_uv[prefix + ".weight_u"], new_uv[prefix + ".weight_v"] = u.clone(), v.clone()
    sigma = torch.dot(u, torch.mv(wm, v))
    return w / sigma


def self_attention(sd: StateDict, p: str, x: Tensor) -> Tensor:
    """vae_laion.py:57-65 -- softmax over the keys of the full (HW x HW) matrix, no 1/sqrt(d) scaling."""
    b, c, h, w = x.shape
    q = F.conv2d(x, sd[p + ".query.weight"], sd[p + ".query.bias"]).view(b, -1, h * w).permute(0, 2, 1)
    k = F.conv2d(x, sd[p + ".key.weight"], sd[p + ".key.bias"]).view(b, -1, h * w)
    v = F.conv2d(x, sd[p + ".value.weight"], sd[p + ".value.bias"]).view(b, -1, h * w)
    out = torch.empty_like(v)
    for i in range(b):                                  # per sample: the matrix is (HW)^2 floats (1 GiB at 128 x 128)
        attn = F.softmax(torch.mm(q[i], k[i]), dim=-1)
        out[i] = torch.mm(v[i], attn.t())
    return sd[p + ".gamma"] * out.view(b, c, h, w) + x


def _bn_eval(sd: StateDict, p: str, x: Tensor) -> Tensor:
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.1, 1e-5)


def residual_block(sd: StateDict, p: str, x: Tensor) -> Tensor:
    """vae_laion.py:80-85 (eval-mode BatchNorm)."""
    y = F.conv2d(x, spectral_weight(sd, p + ".conv1"), None, padding=1)
    y = F.relu(_bn_eval(sd, p + ".bn1", y))
    y = F.conv2d(y, spectral_weight(sd, p + ".conv2"), None, padding=1)
    y = _bn_eval(sd, p + ".bn2", y)
    return y + x


ENC_ATTN = (True, True, False, False)       # vae_laion.py:103,111 (encoder stages 0, 1)
DEC_ATTN = (True, True, False)              # vae_laion.py:142,150 (decoder stages 0, 1)


@torch.no_grad()
def encode(sd: StateDict, x: Tensor):
    """VAE.encode (vae_laion.py:177-184), eval mode.  x: (B, 3, 256, 256) in [0, 1] -> (mu, logvar) (B, latent)."""
    h = x
    for i in range(4):
        p = f"encoder.{i}"
        h = F.relu(F.conv2d(h, spectral_weight(sd, p + ".0"), sd[p + ".0.bias"], stride=2, padding=1))
        h = residual_block(sd, p + ".2", h)
        if ENC_ATTN[i]:
            h = self_attention(sd, p + ".3", h)
    h = h.reshape(h.shape[0], -1)
    return F.linear(h, sd["fc_mu.weight"], sd["fc_mu.bias"]), F.linear(h, sd["fc_logvar.weight"], sd["fc_logvar.bias"])


def reparameterize(mu: Tensor, logvar: Tensor, eps: Tensor) -> Tensor:
    """vae_laion.py:186-189 with the noise injected."""
    return mu + eps * torch.exp(0.5 * logvar)


@torch.no_grad()
def decode(sd: StateDict, z: Tensor) -> Tensor:
    """VAE.decode (vae_laion.py:191-196), eval mode.  z: (B, latent) -> (B, 3, 256, 256) in [0, 1]."""
    h = F.linear(z, sd["decoder_input.weight"], sd["decoder_input.bias"]).view(z.shape[0], 256, 16, 16)
    for i in range(4):
        p = f"decoder.{i}"
        h = F.conv_transpose2d(h, spectral_weight(sd, p + ".0", dim=1), sd[p + ".0.bias"], stride=2, padding=1)
        if i == 3:
            return torch.sigmoid(h)
        h = F.relu(h)
        h = residual_block(sd, p + ".2", h)
        if DEC_ATTN[i]:
            h = self_attention(sd, p + ".3", h)
    return h
