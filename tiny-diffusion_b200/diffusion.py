"""Drop-in for the reference ``diffusion.py`` hot path: ``NoiseModel`` (diffusion.py:11-162),
``ForwardProcess`` (:165-190) and ``sample`` (:254-276), same names, signatures and
``state_dict`` layout, running on hand-written sm_100a kernels (libtinydiff.so).

    from tinydiff.diffusion import NoiseModel, ForwardProcess, sample
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib as L
from .checkpoint import CheckpointCompat
import os

from .process import ForwardProcess, ReverseLoop, SamplerChains
from .unet import MNIST_UNET, UNetConfig, UNetEngine

__all__ = ["NoiseModel", "ForwardProcess", "sample"]


def _cbr(cin: int, cout: int):
    return [nn.Conv2d(cin, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.ReLU()]


class ConvUNetBase(CheckpointCompat, nn.Module):
    """Parameter container + dispatch shared by the three conv UNets.

    The ``torch.nn`` sub-modules exist only to own the parameters/buffers (so ``state_dict()``,
    ``load_state_dict(strict=True)``, ``.parameters()``, ``.to()``, ``.train()/.eval()`` and the
    default initialisers behave exactly like the reference); they are never called.
    """
    config: UNetConfig = MNIST_UNET

    def _build(self, cfg: UNetConfig, time_dim: int, num_classes: Optional[int]):
        c0, c1, c2, c3 = cfg.enc
        d3, d2, d1 = cfg.dec
        self.time_dim = time_dim
        if cfg.cond == "text":
            self.time_mlp = nn.Sequential(nn.Linear(time_dim, time_dim), nn.SiLU(), nn.Linear(time_dim, time_dim))
        else:
            self.time_embedding = nn.Sequential(nn.Linear(1, time_dim), nn.SiLU(), nn.Linear(time_dim, time_dim))
        if cfg.cond == "class":
            self.class_embedding = nn.Embedding(num_classes, time_dim)
        self.initial_conv = nn.Conv2d(cfg.in_ch, c0, 3, padding=1)
        self.enc1 = nn.Sequential(*_cbr(c0, c1), *_cbr(c1, c1))
        self.enc2 = nn.Sequential(*_cbr(c1, c2), *_cbr(c2, c2))
        self.enc3 = nn.Sequential(*_cbr(c2, c3), *_cbr(c3, c3))
        self.bottleneck = nn.Sequential(*_cbr(c3, cfg.bott))
        self.dec3 = nn.Sequential(*_cbr(cfg.bott + c3, d3), *_cbr(d3, d3))
        self.dec2 = nn.Sequential(*_cbr(d3 + c2, d2), *_cbr(d2, d2))
        self.dec1 = nn.Sequential(*_cbr(d2 + c1, d1), *_cbr(d1, d1))
        self.final_conv = nn.Conv2d(d1, cfg.in_ch, 3, padding=1)
        if cfg.cond != "text":       # registration order of the reference (diffusion.py:101-107)
            self.pool = nn.MaxPool2d(2, ceil_mode=cfg.ceil_pool)
            self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
        self.time_proj1 = nn.Conv2d(time_dim, c1, 1)
        self.time_proj2 = nn.Conv2d(time_dim, c2, 1)
        self.time_proj3 = nn.Conv2d(time_dim, c3, 1)
        if cfg.cond == "text":       # conditional_diffusion_laion.py:300-302
            self.pool = nn.MaxPool2d(2)
            self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
        self.precision = "bf16"      # "bf16" (tcgen05) or "fp32" (FFMA parity path)
        self._engines: Dict[Tuple[int, str, str], UNetEngine] = {}

    # engines are not part of the module state
    def __getstate__(self):
        st = self.__dict__.copy()
        st["_engines"] = {}
        st.pop("_sampler_plans", None)
        return st

    def _apply(self, fn, *a, **k):
        self._engines = {}
        return super()._apply(fn, *a, **k)

    def engine(self, batch: int, device: torch.device, replica: int = 0) -> UNetEngine:
        """The eval plan for ``batch`` samples; ``replica`` > 0: another plan of the same shape with its own buffers (the
        sampler runs sub-batches concurrently, process.SamplerChains)."""
        key = (batch, str(device), self.precision) if replica == 0 else (batch, str(device), self.precision, replica)
        eng = self._engines.get(key)
        if eng is None:
            cfg = self.config
            if cfg.time_dim != self.time_dim:
                cfg = UNetConfig(**{**cfg.__dict__, "time_dim": self.time_dim})
            eng = UNetEngine(cfg, self, batch, device, self.precision)
            self._engines[key] = eng
        return eng

    def _forward_impl(self, x, t, cond):
        device = L.require_device(x.device)
        p = next(self.parameters())
        if p.device != device:
            raise RuntimeError(f"NoiseModel parameters are on {p.device}, input on {device}")
        if self.training:
            from .train import unet_train_forward       # batch-statistics BatchNorm (+ autograd)
            return unet_train_forward(self, x, t, cond)
        eng = self.engine(x.shape[0], device)
        return eng.forward(x.to(torch.float32).contiguous(), t, cond)


class NoiseModel(ConvUNetBase):
    """UNet to predict the noise given x_t and t (diffusion.py:11-162)."""
    config = MNIST_UNET

    def __init__(self, time_dim: int = 256):
        super().__init__()
        self._build(self.config, time_dim, None)

    def forward(self, x, t):
        return self._forward_impl(x, t, None)


def sampler_chains(n: int) -> int:
    """Sub-batches one sample() call is split into (TD_SAMPLE_CHAINS; default 1).  Two / four chains at batch 128 measured
    SLOWER (514.7 / 610.0 us per reverse step against 429.3: the persistent convolutions hold every SM, so a second chain
    cannot fill the first one's tails, and every launch's fixed cost doubles; profiles/r02_chains.txt)."""
    want = int(os.environ.get("TD_SAMPLE_CHAINS", "1"))
    if want <= 1 or n < 64:
        return 1
    while want > 1 and n % want:
        want -= 1
    return want


class SamplerPlan:
    """Everything one ``sample()`` call of ``n`` samples needs on the device: the eval plans of its chains (sub-batches), their
    conditioning, and the captured reverse loop.  Cached on the model per (n, process, use_graph)."""

    def __init__(self, noise_model: "ConvUNetBase", diffusion: ForwardProcess, device, n: int, use_graph: bool = True):
        self.model, self.p, self.n, self.device = noise_model, diffusion, n, device
        k = sampler_chains(n)
        self.bounds = [(n * c // k, n * (c + 1) // k) for c in range(k)]
        self.engs = [noise_model.engine(hi - lo, device, replica=c) for c, (lo, hi) in enumerate(self.bounds)]
        self.loop = SamplerChains(diffusion, [{"x": e.x_in, "eps": e.eps, "t_dev": e.t_dev, "launch": e.launch} for e in self.engs],
                                  use_graph=use_graph)
        self.use_graph = use_graph

    def load(self, x_T: torch.Tensor, cond=None) -> None:
        """x_T (any device, [n, C, H, W]) and the conditioning (labels / text embeddings) into the chains' plans."""
        for e, (lo, hi) in zip(self.engs, self.bounds):
            e.refresh_weights()
            e.x_in.copy_(x_T[lo:hi].to(torch.float32), non_blocking=True)
            if e.cfg.cond == "class":
                e.y_in.copy_(cond[lo:hi], non_blocking=True)
            elif e.cfg.cond == "text":
                e.text_in.copy_(cond[lo:hi], non_blocking=True)
            e.use_t_dev = True
            e.prepare_sampler_embed()

    def run(self, z=None, seed: int = 0, steps: Optional[int] = None) -> None:
        self.loop.run(z=z, seed=seed, steps=steps)

    def result(self) -> torch.Tensor:
        out = torch.cat([e.x_in for e in self.engs], dim=0) if len(self.engs) > 1 else self.engs[0].x_in.clone()
        for e in self.engs:
            e.use_t_dev = False
        return out

    def conv_flops(self) -> float:
        return sum(e.conv_flops() for e in self.engs)


def sampler_plan(noise_model, diffusion, device, n: int, use_graph: bool = True) -> SamplerPlan:
    cache = noise_model.__dict__.setdefault("_sampler_plans", {})
    key = (n, str(device), noise_model.precision, id(diffusion), use_graph, sampler_chains(n), id(noise_model._engines))
    plan = cache.get(key)
    if plan is None:
        plan = SamplerPlan(noise_model, diffusion, device, n, use_graph)
        cache.clear()                 # one plan at a time (its graphs pin the chains' buffers)
        cache[key] = plan
    return plan


def _sample_impl(noise_model: ConvUNetBase, diffusion: ForwardProcess, device, shape, cond=None,
                 x_T: Optional[torch.Tensor] = None, z: Optional[torch.Tensor] = None, seed: Optional[int] = None,
                 use_graph: bool = True, steps: Optional[int] = None) -> torch.Tensor:
    device = L.require_device(device)
    noise_model.eval()                                    # diffusion.py:256 (and it stays in eval mode)
    n = shape[0]
    if x_T is None:
        x_T = torch.randn(*shape)                         # CPU generator, then H2D  (diffusion.py:257)
    plan = sampler_plan(noise_model, diffusion, device, n, use_graph)
    plan.load(x_T, cond)
    if z is not None:
        z = z.to(device=device, dtype=torch.float32).contiguous()
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())   # one draw from the CPU generator
    plan.run(z=z, seed=seed, steps=steps)
    return plan.result()


@torch.no_grad()
def sample(noise_model: NoiseModel, diffusion: ForwardProcess, device, n_samples=16, *, x_T=None, z=None,
           seed=None, use_graph=True):
    """diffusion.py:254-276.  Keyword-only extras (not in the reference): ``x_T`` / ``z`` inject the
    initial noise and the per-step noise table [T, n, 1, 28, 28] (parity tests); ``seed`` keys the
    in-kernel Philox stream used otherwise."""
    return _sample_impl(noise_model, diffusion, device, (n_samples, 1, 28, 28), None, x_T, z, seed, use_graph)
