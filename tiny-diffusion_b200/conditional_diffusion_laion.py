"""Drop-in for the reference ``conditional_diffusion_laion.py`` hot path: text-conditioned latent
``NoiseModel`` (conditional_diffusion_laion.py:234-332), ``get_timestep_embedding`` (:223-232),
``ForwardProcess`` (:335-358) and ``sample`` (:560-599).

The denoiser works on (B,4,32,32) latents + a (B,768) text embedding.  The reference's edges --
``diffusers.AutoencoderKL`` and the CLIP text encoder -- need pretrained weights and are not part
of this library: ``sample`` returns the decoded images when a ``vae`` object
with the AutoencoderKL ``decode(...).sample`` interface is passed, and the raw latents otherwise.
"""
from __future__ import annotations

import torch

from . import _lib as L
from .diffusion import ConvUNetBase, _sample_impl
from .process import ForwardProcess
from .unet import LAION_UNET

__all__ = ["NoiseModel", "ForwardProcess", "get_timestep_embedding", "sample"]


def get_timestep_embedding(timesteps: torch.Tensor, embedding_dim: int) -> torch.Tensor:
    """conditional_diffusion_laion.py:223-232 ([sin | cos], divisor half_dim - 1) via td_time_features."""
    import ctypes  # noqa: F401
    device = L.require_device(timesteps.device)
    B = timesteps.shape[0]
    out = torch.zeros(B, embedding_dim, device=device, dtype=torch.float32)
    t = timesteps.to(torch.int64).contiguous()
    L.check(L.load().td_time_features(t.data_ptr(), None, out.data_ptr(), B, embedding_dim, 2, L.stream_ptr()),
            "td_time_features")
    return out


class NoiseModel(ConvUNetBase):
    config = LAION_UNET

    def __init__(self, time_dim: int = 768):
        super().__init__()
        self._build(self.config, time_dim, None)

    def forward(self, x, t, text_embeds):
        return self._forward_impl(x, t, text_embeds)


@torch.no_grad()
def sample(noise_model: NoiseModel, diffusion: ForwardProcess, device, text_embeds=None, vae=None, scaling_factor=1.0,
           *, x_T=None, z=None, seed=None, use_graph=True):
    """conditional_diffusion_laion.py:560-599."""
    if text_embeds is None:
        raise ValueError("Text embeddings must be provided for conditional generation.")
    n = text_embeds.shape[0]
    x = _sample_impl(noise_model, diffusion, device, (n, 4, 32, 32), text_embeds, x_T, z, seed, use_graph)
    if vae is None:
        return x
    decoded = vae.decode(x / scaling_factor).sample               # :589 (external AutoencoderKL)
    images = (decoded / 2 + 0.5).clamp(0, 1)
    bad = torch.logical_or(torch.isnan(images), torch.isinf(images))
    images = torch.where(bad, torch.zeros_like(images), images)   # :591-597
    return images.to(torch.float32)
