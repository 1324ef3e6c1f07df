"""Drop-in for the reference ``diffusion_transformer.py`` hot path: the "DiT" ``NoiseModel``
(diffusion_transformer.py:38-109) with its ``TransformerBlock`` (:16-35), ``ForwardProcess``
(:112-135) and ``sample`` (:291-330).

The reference feeds ``x.unsqueeze(0)`` -- a sequence of length ONE -- to ``nn.MultiheadAttention``
(:99, batch_first=False), so the softmax over a single key is exactly 1 and attention reduces to
``out_proj(V_proj(x))``; the Q/K projections receive an exactly-zero gradient.  The
block is therefore a GEMM chain + LayerNorm, which is what runs here.  Dropout (p = 0.05 by default)
is active in train mode like the reference: on the attention weights (one draw per (sample, head)),
on the FF output (twice, :26 and :34) and on the attention branch (:32), from an in-kernel Philox
stream (statistically, not bitwise, equal to torch's).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib as L
from .dense import DenseEngine, DenseNoiseModel, dense_sample
from .process import ForwardProcess

__all__ = ["NoiseModel", "TransformerBlock", "ForwardProcess", "sample"]


class TransformerBlock(nn.Module):
    """Parameter container with the reference's registration order (diffusion_transformer.py:17-29)."""

    def __init__(self, dim, num_heads, ff_dim, dropout=0.1):
        super().__init__()
        self.attention = nn.MultiheadAttention(dim, num_heads, dropout=dropout)
        self.norm1 = nn.LayerNorm(dim)
        self.ff = nn.Sequential(nn.Linear(dim, ff_dim), nn.GELU(), nn.Linear(ff_dim, dim), nn.Dropout(dropout))
        self.norm2 = nn.LayerNorm(dim)
        self.dropout = nn.Dropout(dropout)
        self.num_heads, self.p = num_heads, float(dropout)


class NoiseModel(DenseNoiseModel):
    emb_mode = 1            # t / 1000 (diffusion_transformer.py:87)

    def __init__(self, time_dim=256, num_classes=10, latent_dim=20, num_heads=4, num_layers=4, dropout=0.05):
        super().__init__()
        self.time_dim, self.latent_dim, self.in_dim = time_dim, latent_dim, latent_dim
        self.time_embedding = nn.Sequential(nn.Linear(1, time_dim), nn.SiLU(), nn.Linear(time_dim, time_dim))
        self.class_embedding = nn.Embedding(num_classes, time_dim)
        self.input_proj = nn.Linear(latent_dim, time_dim)
        self.pos_encoding = nn.Parameter(torch.randn(1, 1, time_dim))
        self.transformer_blocks = nn.ModuleList(
            [TransformerBlock(time_dim, num_heads, time_dim * 4, dropout) for _ in range(num_layers)])
        self.final_layer = nn.Sequential(nn.LayerNorm(time_dim), nn.Linear(time_dim, latent_dim))
        self._init_engines()

    def _declare(self, e: DenseEngine) -> None:
        D = self.time_dim
        F = e.full
        for name, w in (("tfeat", 1), ("h_pre", D), ("h", D), ("emb", D), ("x.0", D)):
            e.new(name, w)
        te = self.time_embedding
        e.time_features(F("tfeat"))
        e.linear("time_embedding.0", F("tfeat"), te[0].weight, te[0].bias, F("h"), act=L.ACT_SILU, pre=F("h_pre"),
                 x_needs_grad=False)
        e.linear("time_embedding.2", F("h"), te[2].weight, te[2].bias, F("emb"), gather=(e.y_in, self.class_embedding.weight))
        # x = input_proj(x) + emb + pos_encoding   (:93-99; the positional "encoding" of a length-1 sequence is a bias)
        e.linear("input_proj", F("x_in"), self.input_proj.weight, self.input_proj.bias, F("x.0"), residual=F("emb"),
                 gather=(e.zero_idx, self.pos_encoding), x_needs_grad=False)
        cur = "x.0"
        for i, blk in enumerate(self.transformer_blocks):
            p = blk.p if e.training else 0.0
            n = lambda s: f"blk{i}.{s}"
            for name, w in ((n("v"), D), (n("s1"), D), (n("x1"), D), (n("u_pre"), 4 * D), (n("u"), 4 * D), (n("s2"), D),
                            (n("x2"), D)):
                e.new(name, w)
            at = blk.attention
            e.linear(n("attn.v_proj"), F(cur), at.in_proj_weight, at.in_proj_bias, F(n("v")), rows=(2 * D, 3 * D))
            if p > 0.0:
                for name in (n("vd"), n("a"), n("ad"), n("f"), n("fd"), n("fdd")):
                    e.new(name, D)
                e.dropout(n("attn.dropout"), F(n("v")), F(n("vd")), p, group=D // blk.num_heads)
                e.linear(n("attn.out_proj"), F(n("vd")), at.out_proj.weight, at.out_proj.bias, F(n("a")))
                e.dropout(n("dropout.attn"), F(n("a")), F(n("ad")), p)
                e.add_into(n("res1.x"), F(cur), F(n("s1")), False)
                e.add_into(n("res1.a"), F(n("ad")), F(n("s1")), True)
            else:
                e.linear(n("attn.out_proj"), F(n("v")), at.out_proj.weight, at.out_proj.bias, F(n("s1")), residual=F(cur))
            e.layernorm(n("norm1"), F(n("s1")), blk.norm1, F(n("x1")))
            e.linear(n("ff.0"), F(n("x1")), blk.ff[0].weight, blk.ff[0].bias, F(n("u")), act=L.ACT_GELU, pre=F(n("u_pre")))
            if p > 0.0:
                e.linear(n("ff.2"), F(n("u")), blk.ff[2].weight, blk.ff[2].bias, F(n("f")))
                e.dropout(n("ff.3"), F(n("f")), F(n("fd")), p)
                e.dropout(n("dropout.ff"), F(n("fd")), F(n("fdd")), p)
                e.add_into(n("res2.x"), F(n("x1")), F(n("s2")), False)
                e.add_into(n("res2.f"), F(n("fdd")), F(n("s2")), True)
            else:
                e.linear(n("ff.2"), F(n("u")), blk.ff[2].weight, blk.ff[2].bias, F(n("s2")), residual=F(n("x1")))
            e.layernorm(n("norm2"), F(n("s2")), blk.norm2, F(n("x2")))
            cur = n("x2")
        e.new("xf", D)
        e.layernorm("final_layer.0", F(cur), self.final_layer[0], F("xf"))
        e.linear("final_layer.1", F("xf"), self.final_layer[1].weight, self.final_layer[1].bias, F("eps"))


@torch.no_grad()
def sample(vae, noise_model: NoiseModel, diffusion: ForwardProcess, device, n_samples=16, y=None, *, x_T=None, z=None,
           seed=None, use_graph=True):
    """diffusion_transformer.py:291-330."""
    return dense_sample(vae, noise_model, diffusion, device, n_samples, y, x_T, z, seed, use_graph)
