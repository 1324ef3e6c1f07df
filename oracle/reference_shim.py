"""Import the UNMODIFIED reference scripts from /root/reference on CPU.  TEST INFRASTRUCTURE ONLY.

The reference cannot be imported as-is offline: every script imports
matplotlib (not installed; e.g. diffusion.py:8), and ``vae.py:87-93`` downloads MNIST at import
time.  The shims below live purely in ``sys.modules`` / monkey-patches; no reference file is
edited or copied.  /root/reference exists only in the build container, so callers must check
``available()`` first; nothing that runs on the GPU box may depend on this module.
"""
from __future__ import annotations

import importlib
import os
import sys
import tempfile
import types

REF_DIR = os.environ.get("TINYDIFF_REFERENCE_DIR", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "diffusion.py"))


def _install_stubs():
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            m = types.ModuleType("matplotlib")
            m.use = lambda *a, **k: None
            sys.modules["matplotlib"] = m
            sys.modules["matplotlib.pyplot"] = types.ModuleType("matplotlib.pyplot")
    if "diffusers" not in sys.modules:
        try:
            import diffusers  # noqa: F401
        except Exception:
            d = types.ModuleType("diffusers")
            d.AutoencoderKL = type("AutoencoderKL", (), {})
            sys.modules["diffusers"] = d


class _FakeMNIST:
    """Stands in for torchvision.datasets.MNIST while ``vae.py`` is imported (vae.py:87-93)."""

    def __init__(self, *a, **k):
        import torch
        self.data = torch.zeros(8, 1, 28, 28)

    def __len__(self):
        return 8

    def __getitem__(self, i):
        return self.data[i], 0


def load(name: str):
    """Return the reference module ``name`` (e.g. "diffusion"), importing it on first use.
    Import happens from a scratch cwd because ``vae.py:101`` creates ./checkpoints."""
    if not available():
        raise RuntimeError(f"reference not present at {REF_DIR}")
    key = f"_tdref_{name}"
    if key in sys.modules:
        return sys.modules[key]
    _install_stubs()
    import torch
    import torchvision
    rng_state = torch.get_rng_state()
    cwd = os.getcwd()
    real_mnist = torchvision.datasets.MNIST
    sys.path.insert(0, REF_DIR)
    try:
        os.chdir(tempfile.mkdtemp(prefix="tdref_"))
        torchvision.datasets.MNIST = _FakeMNIST
        mod = importlib.import_module(name)
    finally:
        torchvision.datasets.MNIST = real_mnist
        os.chdir(cwd)
        sys.path.remove(REF_DIR)
        torch.set_rng_state(rng_state)     # vae.py:33 calls torch.manual_seed(42) at import
    sys.modules[key] = mod
    return mod
