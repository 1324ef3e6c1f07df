"""the vae_laion oracle against the UNMODIFIED reference class (build container only) and against the
reference-generated golden outputs (every box).  CPU only."""
import pytest
import torch

from oracle import reference_shim as shim
from oracle import vae_laion_oracle as V
from oracle.fixtures import checksum, init_state_dict


def _inputs():
    x = torch.rand(1, 3, 256, 256, generator=torch.Generator().manual_seed(77))
    z = torch.randn(1, 128, generator=torch.Generator().manual_seed(78))
    return x, z


def test_oracle_vs_golden(golden):
    g = golden("vae_laion")
    sd = init_state_dict("vae_laion")
    x, z = _inputs()
    mu, logvar = V.encode(sd, x)
    rec = V.decode(sd, z)
    assert torch.allclose(mu, g["mu"], atol=1e-6, rtol=1e-5) and torch.allclose(logvar, g["logvar"], atol=1e-6, rtol=1e-5)
    assert torch.allclose(rec[:, :, ::8, ::8], g["recon_sub"], atol=1e-6)
    assert torch.allclose(checksum(rec), g["recon_checksum"], rtol=1e-6)
    # spectral_norm with one power iteration (training-mode forward of torch's hook)
    pi = g["power_iter"]
    sd2 = dict(sd)
    sd2["encoder.1.0.weight_u"], sd2["encoder.1.0.weight_v"] = pi["u0"], pi["v0"]
    new = {}
    w = V.spectral_weight(sd2, "encoder.1.0", training=True, new_uv=new)
    assert torch.allclose(new["encoder.1.0.weight_u"], pi["u1"], atol=1e-6)
    assert torch.allclose(new["encoder.1.0.weight_v"], pi["v1"], atol=1e-6)
    assert torch.allclose(w, sd["encoder.1.0.weight_orig"] / pi["sigma"], rtol=1e-5, atol=1e-7)


@pytest.mark.skipif(not shim.available(), reason="reference not mounted")
def test_oracle_vs_reference_bitwise():
    mod = shim.load_vae_laion()
    sd = init_state_dict("vae_laion")
    model = shim.build_vae_laion(mod, sd)
    assert [k for k in model.state_dict() if not k.startswith("vgg.")] == list(sd)
    x, z = _inputs()
    with torch.no_grad():
        mu_r, lv_r = model.encode(x)
        rec_r = model.decode(z)
    mu, lv = V.encode(sd, x)
    assert torch.equal(mu, mu_r) and torch.equal(lv, lv_r)
    assert torch.equal(V.decode(sd, z), rec_r)
    # default initialisation: the fixture module reproduces VAE() under the weight seed, key for key
    state = torch.get_rng_state()
    torch.manual_seed(0)
    ref0 = mod.VAE(mod.VAEConfig(device=torch.device("cpu")))
    torch.set_rng_state(state)
    sd0 = init_state_dict("vae_laion", perturb=False)
    assert all(torch.equal(ref0.state_dict()[k], v) for k, v in sd0.items())


def test_dropin_state_dict_layout():
    """tinydiff.vae_laion.VAE: same keys / shapes / default init as the reference (fixture) module."""
    from tinydiff.vae_laion import VAE
    state = torch.get_rng_state()
    torch.manual_seed(0)
    m = VAE()
    torch.set_rng_state(state)
    sd0 = init_state_dict("vae_laion", perturb=False)
    got = m.state_dict()
    assert list(got) == list(sd0)
    assert all(torch.equal(got[k], v) for k, v in sd0.items())
    m.load_state_dict(init_state_dict("vae_laion"), strict=True)
    with pytest.raises(RuntimeError):
        m.eval().encode(torch.zeros(1, 3, 256, 256))            # no CPU fallback
