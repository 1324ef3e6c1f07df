"""Checkpoint compatibility with the reference's ``torch.save(noise_model.state_dict(), path)`` files.

The state_dict layout of every drop-in module equals the reference's, so plain checkpoints load
unchanged.  The LAION script wraps its model in ``torch.compile`` before training
(conditional_diffusion_laion.py:614), so the checkpoints it writes (:527) carry an ``_orig_mod.`` prefix on every key;
``strip_compile_prefix`` removes it and the drop-in modules' ``load_state_dict`` applies it automatically."""
from typing import Dict, Mapping

import torch

_PREFIX = "_orig_mod."


def strip_compile_prefix(state_dict: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    if not any(k.startswith(_PREFIX) for k in state_dict):
        return dict(state_dict)
    return {(k[len(_PREFIX):] if k.startswith(_PREFIX) else k): v for k, v in state_dict.items()}


class CheckpointCompat:
    """Mixin for the drop-in ``nn.Module``s: ``load_state_dict`` accepts the compiled-module key prefix."""

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        out = super().load_state_dict(strip_compile_prefix(state_dict), strict=strict, assign=assign)
        if assign and hasattr(self, "_engines"):
            # assign=True swaps the parameter / buffer storages: the plans cache raw device pointers of the old ones
            self._engines = {}
        self._weights_gen = getattr(self, "_weights_gen", 0) + 1
        return out


def load_checkpoint(model: torch.nn.Module, path: str, map_location="cpu", strict: bool = True):
    """``model.load_state_dict(torch.load(path))`` for files written by any of the reference scripts."""
    sd = torch.load(path, map_location=map_location, weights_only=True)
    return model.load_state_dict(strip_compile_prefix(sd), strict=strict)
