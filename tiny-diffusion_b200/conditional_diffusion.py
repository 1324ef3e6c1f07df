"""Drop-in for the reference ``conditional_diffusion.py`` hot path: class-conditional
``NoiseModel`` (conditional_diffusion.py:14-171), ``ForwardProcess`` (:174-199) and ``sample``
(:354-386)."""
from __future__ import annotations

import torch

from .diffusion import ConvUNetBase, _sample_impl
from .process import ForwardProcess
from .unet import COND_UNET

__all__ = ["NoiseModel", "ForwardProcess", "sample"]


class NoiseModel(ConvUNetBase):
    """UNet to predict the noise given x_t, t, and class label y."""
    config = COND_UNET

    def __init__(self, time_dim: int = 256, num_classes: int = 10):
        super().__init__()
        self._build(self.config, time_dim, num_classes)

    def forward(self, x, t, y):
        return self._forward_impl(x, t, y)


@torch.no_grad()
def sample(noise_model: NoiseModel, diffusion: ForwardProcess, device, n_samples=16, y=None, *, x_T=None, z=None,
           seed=None, use_graph=True):
    """conditional_diffusion.py:354-386 (plain conditional ancestral sampling; the reference has no
    classifier-free guidance)."""
    if y is None:
        raise ValueError("Class labels 'y' must be provided for conditional generation.")
    if y.shape[0] != n_samples:
        raise ValueError("y must have shape (n_samples,)")
    return _sample_impl(noise_model, diffusion, device, (n_samples, 1, 28, 28), y, x_T, z, seed, use_graph)
