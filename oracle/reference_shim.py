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


def load_vae_laion():
    """Import the reference's ``vae_laion.py`` on CPU and return the module (its ``VAE`` / ``VAEConfig`` / ``SelfAttention``
    / ``ResidualBlock`` classes are the parity target of SURVEY.md 8f #3).  The script cannot be imported offline as-is:
    at import time it loads a HuggingFace dataset (vae_laion.py:322), builds the model on a hard-coded
    ``torch.device("cuda")`` (:34,:335) and every ``VAE()`` downloads VGG16 weights (:172).  Stubs, in ``sys.modules`` /
    monkey-patches only, active during the import: ``datasets.load_dataset`` -> a one-row list,
    ``torchvision.models.vgg16`` -> a tiny random feature stack (only the perceptual loss of the TRAINING objective uses it;
    encode / decode never do), ``nn.Module.to`` / ``Tensor.to`` with a CUDA device -> no-op."""
    if not available():
        raise RuntimeError(f"reference not present at {REF_DIR}")
    key = "_tdref_vae_laion"
    if key in sys.modules:
        return sys.modules[key]
    _install_stubs()
    import torch
    import torch.nn as nn
    import torchvision
    import datasets as hf_datasets

    class _FakeVGG(nn.Module):
        def __init__(self):
            super().__init__()
            self.features = nn.Sequential(*[nn.Identity() for _ in range(31)])

    def _is_cuda(a):
        return isinstance(a, torch.device) and a.type == "cuda" or (isinstance(a, str) and a.startswith("cuda"))

    real_to, real_load, real_vgg = nn.Module.to, hf_datasets.load_dataset, torchvision.models.vgg16

    def _to(self, *a, **k):
        if any(_is_cuda(x) for x in a) or _is_cuda(k.get("device")):
            return self
        return real_to(self, *a, **k)

    rng_state = torch.get_rng_state()
    cwd = os.getcwd()
    sys.path.insert(0, REF_DIR)
    try:
        os.chdir(tempfile.mkdtemp(prefix="tdref_"))
        nn.Module.to = _to
        hf_datasets.load_dataset = lambda *a, **k: [{"URL": "", "TEXT": ""}]
        torchvision.models.vgg16 = lambda *a, **k: _FakeVGG()
        mod = importlib.import_module("vae_laion")
    finally:
        nn.Module.to = real_to
        hf_datasets.load_dataset = real_load
        torchvision.models.vgg16 = real_vgg
        os.chdir(cwd)
        sys.path.remove(REF_DIR)
        torch.set_rng_state(rng_state)     # vae_laion.py:46 calls torch.manual_seed(42) at import
    # VAE() looks `vgg16` up in the module's own globals (bound by `from torchvision.models import vgg16` at import): keep
    # the stub there, and let `.to(config.device)` inside VAE.__init__ (:173) be a no-op for a CPU config
    mod.vgg16 = lambda *a, **k: _FakeVGG()
    sys.modules[key] = mod
    return mod


def build_vae_laion(mod, state_dict):
    """The reference's own ``VAE`` on CPU carrying ``state_dict`` (the fixture weights; VGG stub keys excluded)."""
    import torch
    cfg = mod.VAEConfig(device=torch.device("cpu"))
    model = mod.VAE(cfg)
    missing, unexpected = model.load_state_dict(state_dict, strict=False)
    assert not unexpected and all(k.startswith("vgg.") for k in missing), (missing, unexpected)
    return model.eval()
