"""Seeded weights / inputs shared by the golden generator and the tests.  TEST INFRASTRUCTURE ONLY.

``init_state_dict`` restates the reference ``NoiseModel.__init__`` bodies (registration order
matters: it fixes both the ``state_dict`` key order and the order in which the default
initialisers consume the global RNG), so that ``torch.manual_seed(0)`` followed by
``init_state_dict(name)`` yields bit-identical weights to ``reference.NoiseModel()`` under the
same seed -- which ``tests/test_oracle_vs_reference.py`` checks.
"""
from __future__ import annotations

import contextlib
from typing import Dict, List

import torch
import torch.nn as nn

WEIGHT_SEED = 0
DATA_SEED = 1234
BN_SEED = 4321


def _cbr(cin, cout):
    return [nn.Conv2d(cin, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.ReLU()]


def _unet_modules(in_ch, c0, c1, c2, c3, cb, time_dim, cond: str) -> nn.Module:
    """diffusion.py:16-107 / conditional_diffusion.py:19-110 / conditional_diffusion_laion.py:235-302."""
    m = nn.Module()
    if cond == "sinusoidal+text":
        m.time_mlp = nn.Sequential(nn.Linear(time_dim, time_dim), nn.SiLU(), nn.Linear(time_dim, time_dim))
    else:
        m.time_embedding = nn.Sequential(nn.Linear(1, time_dim), nn.SiLU(), nn.Linear(time_dim, time_dim))
    if cond == "linear+class":
        m.class_embedding = nn.Embedding(10, time_dim)
    m.initial_conv = nn.Conv2d(in_ch, c0, 3, padding=1)
    m.enc1 = nn.Sequential(*_cbr(c0, c1), *_cbr(c1, c1))
    m.enc2 = nn.Sequential(*_cbr(c1, c2), *_cbr(c2, c2))
    m.enc3 = nn.Sequential(*_cbr(c2, c3), *_cbr(c3, c3))
    m.bottleneck = nn.Sequential(*_cbr(c3, cb))
    m.dec3 = nn.Sequential(*_cbr(cb + c3, c2 if cond != "sinusoidal+text" else c3),
                           *_cbr(c2 if cond != "sinusoidal+text" else c3,
                                 c2 if cond != "sinusoidal+text" else c3))
    if cond != "sinusoidal+text":
        m.dec2 = nn.Sequential(*_cbr(c2 + c2, c1), *_cbr(c1, c1))
        m.dec1 = nn.Sequential(*_cbr(c1 + c1, c0), *_cbr(c0, c0))
        m.final_conv = nn.Conv2d(c0, in_ch, 3, padding=1)
    else:
        m.dec2 = nn.Sequential(*_cbr(c3 + c2, c2), *_cbr(c2, c2))
        m.dec1 = nn.Sequential(*_cbr(c2 + c1, c1), *_cbr(c1, c1))
        m.final_conv = nn.Conv2d(c1, in_ch, 3, padding=1)
    m.time_proj1 = nn.Conv2d(time_dim, c1, 1)
    m.time_proj2 = nn.Conv2d(time_dim, c2, 1)
    m.time_proj3 = nn.Conv2d(time_dim, c3, 1)
    return m


class _Block(nn.Module):
    """diffusion_transformer.py:16-29 (parameter containers only)."""

    def __init__(self, dim, heads, ff, dropout):
        super().__init__()
        self.attention = nn.MultiheadAttention(dim, heads, dropout=dropout)
        self.norm1 = nn.LayerNorm(dim)
        self.ff = nn.Sequential(nn.Linear(dim, ff), nn.GELU(), nn.Linear(ff, dim), nn.Dropout(dropout))
        self.norm2 = nn.LayerNorm(dim)
        self.dropout = nn.Dropout(dropout)


def _dit_modules(time_dim=256, num_classes=10, latent_dim=20, heads=4, layers=4, dropout=0.0):
    """diffusion_transformer.py:48-79."""
    m = nn.Module()
    m.time_embedding = nn.Sequential(nn.Linear(1, time_dim), nn.SiLU(), nn.Linear(time_dim, time_dim))
    m.class_embedding = nn.Embedding(num_classes, time_dim)
    m.input_proj = nn.Linear(latent_dim, time_dim)
    m.pos_encoding = nn.Parameter(torch.randn(1, 1, time_dim))
    m.transformer_blocks = nn.ModuleList([_Block(time_dim, heads, time_dim * 4, dropout)
                                          for _ in range(layers)])
    m.final_layer = nn.Sequential(nn.LayerNorm(time_dim), nn.Linear(time_dim, latent_dim))
    return m


def _lbr(cin, cout):
    return [nn.Linear(cin, cout), nn.BatchNorm1d(cout), nn.ReLU()]


def _mlp_modules(time_dim=256, num_classes=10, latent_dim=20) -> nn.Module:
    """latent_diffusion.py:17-105 (registration order = RNG consumption order)."""
    m = nn.Module()
    m.time_embedding = nn.Sequential(nn.Linear(1, time_dim), nn.SiLU(), nn.Linear(time_dim, time_dim))
    m.class_embedding = nn.Embedding(num_classes, time_dim)
    m.initial_fc = nn.Linear(latent_dim, 512)
    m.enc1 = nn.Sequential(*_lbr(512, 512), *_lbr(512, 256))
    m.enc2 = nn.Sequential(*_lbr(256, 256), *_lbr(256, 128))
    m.enc3 = nn.Sequential(*_lbr(128, 128), *_lbr(128, 64))
    m.bottleneck = nn.Sequential(*_lbr(64, 64))
    m.dec3 = nn.Sequential(*_lbr(128, 128), *_lbr(128, 128))
    m.dec2 = nn.Sequential(*_lbr(256, 256), *_lbr(256, 256))
    m.dec1 = nn.Sequential(*_lbr(512, 512), *_lbr(512, 512))
    m.final_fc = nn.Linear(512, latent_dim)
    m.time_proj1 = nn.Linear(time_dim, 64)
    m.time_proj2 = nn.Linear(time_dim, 128)
    m.time_proj3 = nn.Linear(time_dim, 256)
    return m


def _vae_modules(input_dim=784, hidden_dim=400, latent_dim=20) -> nn.Module:
    """vae.py:42-49."""
    m = nn.Module()
    m.fc1 = nn.Linear(input_dim, hidden_dim)
    m.fc21 = nn.Linear(hidden_dim, latent_dim)
    m.fc22 = nn.Linear(hidden_dim, latent_dim)
    m.fc3 = nn.Linear(latent_dim, hidden_dim)
    m.fc4 = nn.Linear(hidden_dim, input_dim)
    return m


class _SelfAttentionP(nn.Module):
    """Parameter layout of vae_laion.SelfAttention (vae_laion.py:51-55)."""

    def __init__(self, c):
        super().__init__()
        self.query = nn.Conv2d(c, c // 8, 1)
        self.key = nn.Conv2d(c, c // 8, 1)
        self.value = nn.Conv2d(c, c, 1)
        self.gamma = nn.Parameter(torch.zeros(1))


class _ResidualBlockP(nn.Module):
    """vae_laion.py:70-78."""

    def __init__(self, c):
        super().__init__()
        self.conv1 = nn.utils.spectral_norm(nn.Conv2d(c, c, 3, padding=1, bias=False))
        self.bn1 = nn.BatchNorm2d(c)
        self.conv2 = nn.utils.spectral_norm(nn.Conv2d(c, c, 3, padding=1, bias=False))
        self.bn2 = nn.BatchNorm2d(c)


def _vae_laion_modules(latent_dim=128, in_ch=3) -> nn.Module:
    """vae_laion.VAE.__init__ (vae_laion.py:94-168) without the VGG16 perceptual-loss network (:171-176, pretrained
    weights; its parameters are not part of the encode / decode path).  Same construction order, hence the same
    default-init random stream and the same state_dict keys."""
    sn = nn.utils.spectral_norm
    m = nn.Module()
    m.encoder = nn.ModuleList([
        nn.Sequential(sn(nn.Conv2d(in_ch, 32, 4, stride=2, padding=1)), nn.ReLU(), _ResidualBlockP(32), _SelfAttentionP(32)),
        nn.Sequential(sn(nn.Conv2d(32, 64, 4, stride=2, padding=1)), nn.ReLU(), _ResidualBlockP(64), _SelfAttentionP(64)),
        nn.Sequential(sn(nn.Conv2d(64, 128, 4, stride=2, padding=1)), nn.ReLU(), _ResidualBlockP(128)),
        nn.Sequential(sn(nn.Conv2d(128, 256, 4, stride=2, padding=1)), nn.ReLU(), _ResidualBlockP(256)),
    ])
    m.fc_mu = nn.Linear(256 * 16 * 16, latent_dim)
    m.fc_logvar = nn.Linear(256 * 16 * 16, latent_dim)
    m.decoder_input = nn.Linear(latent_dim, 256 * 16 * 16)
    m.decoder = nn.ModuleList([
        nn.Sequential(sn(nn.ConvTranspose2d(256, 128, 4, stride=2, padding=1)), nn.ReLU(), _ResidualBlockP(128), _SelfAttentionP(128)),
        nn.Sequential(sn(nn.ConvTranspose2d(128, 64, 4, stride=2, padding=1)), nn.ReLU(), _ResidualBlockP(64), _SelfAttentionP(64)),
        nn.Sequential(sn(nn.ConvTranspose2d(64, 32, 4, stride=2, padding=1)), nn.ReLU(), _ResidualBlockP(32)),
        nn.Sequential(sn(nn.ConvTranspose2d(32, in_ch, 4, stride=2, padding=1)), nn.Sigmoid()),
    ])
    return m


def perturb_vae_laion(sd: Dict[str, torch.Tensor], seed: int = BN_SEED) -> Dict[str, torch.Tensor]:
    """The attention gates ``gamma`` are initialised to ZERO (vae_laion.py:55): every SelfAttention would be the identity and
    the attention kernel untested.  Give them O(1) values (and perturb the BatchNorm statistics like ``perturb_bn``)."""
    g = torch.Generator().manual_seed(seed + 1)
    out = dict(sd)
    for k in sd:
        if k.endswith(".gamma"):
            out[k] = 0.5 + torch.rand(1, generator=g)
    # u / v of torch.nn.utils.spectral_norm are random at construction and only become singular vectors through the power
    # iterations of TRAINING-mode forwards; in eval mode sigma = u^T W v of the raw vectors is a random, near-zero number and
    # W / sigma explodes.  Run the power iteration (spectral_norm's own update) a few times, as any trained checkpoint has.
    for k in sd:
        if not k.endswith(".weight_orig"):
            continue
        p = k[:-len(".weight_orig")]
        w = sd[k]
        transposed = p.startswith("decoder.") and p.endswith(".0")          # ConvTranspose2d: dim = 1
        wm = (w.permute(1, 0, 2, 3) if transposed else w).reshape(w.shape[1] if transposed else w.shape[0], -1)
        u, v = sd[p + ".weight_u"].clone(), sd[p + ".weight_v"].clone()
        for _ in range(8):
            v = torch.nn.functional.normalize(torch.mv(wm.t(), u), dim=0, eps=1e-12)
            u = torch.nn.functional.normalize(torch.mv(wm, v), dim=0, eps=1e-12)
        out[p + ".weight_u"], out[p + ".weight_v"] = u, v
    return out


def make_modules(modname: str) -> nn.Module:
    if modname == "diffusion":
        return _unet_modules(1, 64, 128, 256, 512, 512, 256, "linear")
    if modname == "conditional_diffusion":
        return _unet_modules(1, 64, 128, 256, 512, 512, 256, "linear+class")
    if modname == "conditional_diffusion_laion":
        return _unet_modules(4, 32, 64, 128, 256, 256, 768, "sinusoidal+text")
    if modname == "diffusion_transformer":
        return _dit_modules()
    if modname == "latent_diffusion":
        return _mlp_modules()
    if modname == "vae":
        return _vae_modules()
    if modname == "vae_laion":
        return _vae_laion_modules()
    raise KeyError(modname)


def perturb_bn(model: nn.Module, seed: int = BN_SEED) -> nn.Module:
    """Make BatchNorm non-trivial (default init is gamma=1, beta=0, mean=0, var=1)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for mod in model.modules():
            if isinstance(mod, (nn.BatchNorm1d, nn.BatchNorm2d)):
                n = mod.num_features
                mod.weight.copy_(0.5 + torch.rand(n, generator=g))
                mod.bias.copy_(0.1 * torch.randn(n, generator=g))
                mod.running_mean.copy_(0.1 * torch.randn(n, generator=g))
                mod.running_var.copy_(0.5 + torch.rand(n, generator=g))
    return model


def init_state_dict(modname: str, seed: int = WEIGHT_SEED, perturb: bool = True) -> Dict[str, torch.Tensor]:
    state = torch.get_rng_state()
    try:
        torch.manual_seed(seed)
        m = make_modules(modname)
    finally:
        torch.set_rng_state(state)
    if perturb:
        perturb_bn(m)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    if modname == "vae_laion" and perturb:
        sd = perturb_vae_laion(sd)
    return sd


def build_reference_model(ref_module, modname: str, seed: int = WEIGHT_SEED):
    """Reference NoiseModel() under the weight seed (drawn AFTER import: vae.py:33 reseeds)."""
    state = torch.get_rng_state()
    try:
        torch.manual_seed(seed)
        if modname == "diffusion_transformer":
            model = ref_module.NoiseModel(dropout=0.0)
        else:
            model = ref_module.NoiseModel()
    finally:
        torch.set_rng_state(state)
    return model


def make_inputs(modname: str, B: int, seed: int = DATA_SEED) -> Dict[str, torch.Tensor]:
    """Synthetic inputs of SURVEY.md section 8(d)."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    if modname in ("diffusion", "conditional_diffusion"):
        out["x0"] = torch.rand(B, 1, 28, 28, generator=g) * 2 - 1
    elif modname == "conditional_diffusion_laion":
        out["x0"] = 0.18215 * torch.randn(B, 4, 32, 32, generator=g)
    else:
        out["x0"] = torch.randn(B, 20, generator=g)
    out["t"] = torch.randint(0, 1000, (B,), generator=g)
    out["noise"] = torch.randn(out["x0"].shape, generator=g)
    if modname in ("conditional_diffusion", "diffusion_transformer", "latent_diffusion"):
        out["cond"] = torch.randint(0, 10, (B,), generator=g)
    elif modname == "conditional_diffusion_laion":
        out["cond"] = torch.randn(B, 768, generator=g)
    return out


def checksum(x: torch.Tensor) -> torch.Tensor:
    """(sum, L2 norm, a fixed pseudo-random projection) in fp64 -- cheap to store, hard to fake."""
    xd = x.detach().double().flatten()
    n = xd.numel()
    idx = torch.arange(n, dtype=torch.float64)
    proj = torch.cos(idx * 0.61803398875 + 0.25)
    return torch.stack([xd.sum(), xd.norm(), (xd * proj).sum()])


@contextlib.contextmanager
def injected_randn(ref_module, tensors: List[torch.Tensor]):
    """Make the reference module's ``torch.randn`` / ``torch.randn_like`` return the given
    tensors in order (SURVEY.md section 8b RNG contract: q_sample draws once, diffusion.py:178;
    sample() draws x_T then one z per step t>0, diffusion.py:257,268)."""
    queue = list(tensors)
    real = ref_module.torch

    class Proxy:
        def __getattr__(self, name):
            return getattr(real, name)

        @staticmethod
        def randn(*a, **k):
            return queue.pop(0).clone()

        @staticmethod
        def randn_like(x, **k):
            t = queue.pop(0)
            assert t.shape == x.shape, (t.shape, x.shape)
            return t.clone()

    ref_module.torch = Proxy()
    try:
        yield
    finally:
        ref_module.torch = real
