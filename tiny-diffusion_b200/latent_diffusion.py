"""Drop-in for the reference ``latent_diffusion.py`` hot path: the class-conditional MLP "U-Net"
``NoiseModel`` on 20-d VAE latents (latent_diffusion.py:16-128), ``ForwardProcess`` (:131-154) and
``sample`` (:308-347)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib as L
from .dense import DenseEngine, DenseNoiseModel, dense_sample
from .process import ForwardProcess

__all__ = ["NoiseModel", "ForwardProcess", "sample"]


def _lbr(cin: int, cout: int):
    return [nn.Linear(cin, cout), nn.BatchNorm1d(cout), nn.ReLU()]


class NoiseModel(DenseNoiseModel):
    emb_mode = 0            # raw t (latent_diffusion.py:108)

    def __init__(self, time_dim: int = 256, num_classes: int = 10, latent_dim: int = 20):
        super().__init__()
        self.time_dim, self.latent_dim, self.in_dim = time_dim, latent_dim, latent_dim
        self.time_embedding = nn.Sequential(nn.Linear(1, time_dim), nn.SiLU(), nn.Linear(time_dim, time_dim))
        self.class_embedding = nn.Embedding(num_classes, time_dim)
        self.initial_fc = nn.Linear(latent_dim, 512)
        self.enc1 = nn.Sequential(*_lbr(512, 512), *_lbr(512, 256))
        self.enc2 = nn.Sequential(*_lbr(256, 256), *_lbr(256, 128))
        self.enc3 = nn.Sequential(*_lbr(128, 128), *_lbr(128, 64))
        self.bottleneck = nn.Sequential(*_lbr(64, 64))
        self.dec3 = nn.Sequential(*_lbr(128, 128), *_lbr(128, 128))
        self.dec2 = nn.Sequential(*_lbr(256, 256), *_lbr(256, 256))
        self.dec1 = nn.Sequential(*_lbr(512, 512), *_lbr(512, 512))
        self.final_fc = nn.Linear(512, latent_dim)
        self.time_proj1 = nn.Linear(time_dim, 64)
        self.time_proj2 = nn.Linear(time_dim, 128)
        self.time_proj3 = nn.Linear(time_dim, 256)
        self._init_engines()

    def _declare(self, e: DenseEngine) -> None:
        D = self.time_dim
        for name, w in (("tfeat", 1), ("h_pre", D), ("h", D), ("emb", D), ("x0", 512), ("bt", 64), ("d3", 128),
                        ("d2", 256), ("d1", 512), ("cat3", 128), ("cat2", 256), ("cat1", 512)):
            e.new(name, w)
        F = e.full
        te = self.time_embedding
        e.time_features(F("tfeat"))
        e.linear("time_embedding.0", F("tfeat"), te[0].weight, te[0].bias, F("h"), act=L.ACT_SILU, pre=F("h_pre"),
                 x_needs_grad=False)
        e.linear("time_embedding.2", F("h"), te[2].weight, te[2].bias, F("emb"), gather=(e.y_in, self.class_embedding.weight))
        e.linear("initial_fc", F("x_in"), self.initial_fc.weight, self.initial_fc.bias, F("x0"), x_needs_grad=False)

        def lbr(blk: str, seq, idx: int, x, out):
            lin, bn = seq[idx], seq[idx + 1]
            y = e.new(f"{blk}.{idx}:y", lin.out_features) is not None and F(f"{blk}.{idx}:y")
            e.linear(f"{blk}.{idx}", x, lin.weight, lin.bias, y)
            e.bn1d(f"{blk}.{idx + 1}", y, bn, out)
            return out

        def double(blk: str, seq, x, out):
            mid = seq[0].out_features
            e.new(f"{blk}.0:a", mid)
            return lbr(blk, seq, 3, lbr(blk, seq, 0, x, F(f"{blk}.0:a")), out)

        e1, e2, e3 = ("cat1", 256, 512), ("cat2", 128, 256), ("cat3", 64, 128)
        double("enc1", self.enc1, F("x0"), e1)
        double("enc2", self.enc2, e1, e2)
        double("enc3", self.enc3, e2, e3)
        lbr("bottleneck", self.bottleneck, 0, e3, F("bt"))
        # the embedding goes to the TRUNK: cat_k[:, :c] = trunk + time_proj_k(emb)  (latent_diffusion.py:119-125)
        e.linear("time_proj1", F("emb"), self.time_proj1.weight, self.time_proj1.bias, ("cat3", 0, 64), residual=F("bt"))
        double("dec3", self.dec3, F("cat3"), F("d3"))
        e.linear("time_proj2", F("emb"), self.time_proj2.weight, self.time_proj2.bias, ("cat2", 0, 128), residual=F("d3"))
        double("dec2", self.dec2, F("cat2"), F("d2"))
        e.linear("time_proj3", F("emb"), self.time_proj3.weight, self.time_proj3.bias, ("cat1", 0, 256), residual=F("d2"))
        double("dec1", self.dec1, F("cat1"), F("d1"))
        e.linear("final_fc", F("d1"), self.final_fc.weight, self.final_fc.bias, F("eps"))


@torch.no_grad()
def sample(vae, noise_model: NoiseModel, diffusion: ForwardProcess, device, n_samples=16, y=None, *, x_T=None, z=None,
           seed=None, use_graph=True):
    """latent_diffusion.py:308-347: ancestral sampling in latent space, then ``vae.decode(z).view(-1,1,28,28)``."""
    return dense_sample(vae, noise_model, diffusion, device, n_samples, y, x_T, z, seed, use_graph)
