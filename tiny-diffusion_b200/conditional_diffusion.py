"""Drop-in for the reference ``conditional_diffusion.py`` hot path: class-conditional
``NoiseModel`` (conditional_diffusion.py:14-171), ``ForwardProcess`` (:174-199) and ``sample``
(:354-386)."""
from __future__ import annotations

import torch

from .diffusion import ConvUNetBase, _sample_impl
from .process import ForwardProcess
from .unet import COND_UNET

__all__ = ["NoiseModel", "ForwardProcess", "sample", "sample_cfg", "cfg_drop_labels"]


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


# ---------------------------------------------------------------------------------------------
# Classifier-free guidance -- an EXTENSION, off unless called.  BASELINE.json's config 2 names it, the reference does
# not have it (no label dropout, no null class, no guidance scale: SURVEY.md D5), so there is no reference behaviour to
# match; the parity target is the composition eps_u + w*(eps_c - eps_u) of two reference forwards, which
# tests/test_gpu_model.py checks against the oracle.  Convention: the null label is the LAST row of the class embedding,
# i.e. build the model as NoiseModel(num_classes=11) and train with cfg_drop_labels(y, p_uncond=0.1).
# ---------------------------------------------------------------------------------------------
def cfg_drop_labels(y: torch.Tensor, p_uncond: float, null_label: int, generator=None) -> torch.Tensor:
    """Training-time label dropout: each label is replaced by ``null_label`` with probability ``p_uncond``."""
    drop = torch.rand(y.shape, generator=generator, device=y.device if generator is None else generator.device) < p_uncond
    return torch.where(drop.to(y.device), torch.full_like(y, null_label), y)


@torch.no_grad()
def sample_cfg(noise_model: NoiseModel, diffusion: ForwardProcess, device, n_samples=16, y=None, guidance_scale: float = 3.0,
               null_label=None, *, x_T=None, z=None, seed=None, use_graph=True):
    """Ancestral sampling with classifier-free guidance: every reverse step runs ONE denoiser forward on a doubled batch
    (rows [0, n): labels ``y``; rows [n, 2n): ``null_label``) and the fused step kernel ``td_psample_step_cfg`` combines
    ``eps_u + w*(eps_c - eps_u)`` with the x_{t-1} update.  ``guidance_scale = 1`` reproduces ``sample`` up to rounding."""
    if y is None:
        raise ValueError("Class labels 'y' must be provided for conditional generation.")
    if y.shape[0] != n_samples:
        raise ValueError("y must have shape (n_samples,)")
    from . import _lib as L
    from .process import ReverseLoop
    device = L.require_device(device)
    num_classes = noise_model.class_embedding.num_embeddings
    if null_label is None:
        null_label = num_classes - 1
    if not 0 <= int(null_label) < num_classes:
        raise ValueError(f"null_label {null_label} outside the class embedding (num_classes={num_classes})")
    noise_model.eval()
    n = n_samples
    if x_T is None:
        x_T = torch.randn(n, 1, 28, 28)                   # CPU generator, as conditional_diffusion.py:367
    eng = noise_model.engine(2 * n, device)
    eng.refresh_weights()
    x_T = x_T.to(device=device, dtype=torch.float32)
    eng.x_in.copy_(torch.cat([x_T, x_T], dim=0))
    y = y.to(device)
    eng.y_in.copy_(torch.cat([y, torch.full_like(y, int(null_label))], dim=0))
    eng.use_t_dev = True
    eng.prepare_sampler_embed()
    loop = getattr(eng, "_reverse_loop_cfg", None)
    if loop is None or loop.p is not diffusion or loop.use_graph != use_graph or loop.guidance != float(guidance_scale):
        loop = ReverseLoop(diffusion, eng.x_in, eng.eps, eng.t_dev, eng.launch, use_graph=use_graph,
                           guidance=float(guidance_scale))
        eng._reverse_loop_cfg = loop
    if z is not None:
        z = z.to(device=device, dtype=torch.float32).contiguous()
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    loop.run(z=z, seed=seed)
    out = eng.x_in[:n].clone()
    eng.use_t_dev = False
    return out
