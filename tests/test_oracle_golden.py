"""The oracle (oracle/ddpm_oracle.py) against the reference-generated fixtures in tests/golden/.
Runs everywhere (CPU only; /root/reference not needed)."""
import pytest
import torch

from oracle import ddpm_oracle as O
from oracle.fixtures import checksum, init_state_dict, make_inputs

SPECS = {"diffusion": O.UNET_MNIST, "conditional_diffusion": O.UNET_COND,
         "conditional_diffusion_laion": O.UNET_LAION}


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def test_schedule_known_answers(golden):
    """SURVEY.md section 8(c) known-answer anchors + the stored reference tensors (bit-exact)."""
    betas, alphas, ac = O.make_schedule()
    g = golden("diffusion")["schedule"]
    assert torch.equal(betas, g["betas"]) and torch.equal(alphas, g["alphas"])
    assert torch.equal(ac, g["alphas_cumprod"])
    assert float(betas[0]) == pytest.approx(9.999999747e-05, rel=1e-7)
    assert float(ac[-1]) == pytest.approx(4.035830e-05, rel=1e-5)
    c1, c2, c3 = O.p_sample_coeffs(betas, alphas, ac)
    assert float(c1[500]) == pytest.approx(1.0050681829, rel=1e-7)
    assert float(c2[500]) == pytest.approx(0.0104756886, rel=1e-6)
    assert float(c3[500]) == pytest.approx(0.1002993509, rel=1e-7)


@pytest.mark.parametrize("name", ["diffusion", "conditional_diffusion", "conditional_diffusion_laion"])
def test_unet_forward_and_grads(golden, name):
    g = golden(name)
    spec = SPECS[name]
    B = g["x_t"].shape[0]
    sd = init_state_dict(name)
    inp = make_inputs(name, B)
    _, _, ac = O.make_schedule()
    x_t = O.q_sample(ac, inp["x0"], inp["t"], inp["noise"])
    assert torch.equal(x_t, g["x_t"])
    with torch.no_grad():
        eps = O.unet_forward(spec, sd, x_t, inp["t"], inp.get("cond"))
    assert rel(eps, g["eps_eval"]) < 1e-6
    loss, grads, new_stats, pred = O.unet_loss_and_grads(spec, sd, inp["x0"], inp["t"], inp["noise"], ac,
                                                         inp.get("cond"))
    assert rel(pred, g["eps_train"]) < 1e-6
    assert float(loss) == pytest.approx(float(g["loss"]), rel=1e-6)
    for k, cs in g["grad_checksums"].items():
        mine = checksum(grads[k])
        # conv biases in front of a train-mode BN have mathematically zero gradient (pure noise)
        if float(cs[1]) < 1e-6:
            assert float(mine[1]) < 1e-5, k
        else:
            assert float(mine[1]) == pytest.approx(float(cs[1]), rel=1e-4), k
    if "sin_emb" in g:
        assert torch.equal(O.timestep_embedding_sinusoidal(inp["t"], 768), g["sin_emb"])


@pytest.mark.parametrize("name", ["diffusion", "conditional_diffusion"])
def test_adam_step(golden, name):
    g = golden(name)
    spec = SPECS[name]
    B = g["x_t"].shape[0]
    sd = init_state_dict(name)
    inp = make_inputs(name, B)
    _, _, ac = O.make_schedule()
    loss, grads, new_stats, _ = O.unet_loss_and_grads(spec, sd, inp["x0"], inp["t"], inp["noise"], ac,
                                                      inp.get("cond"))
    for k, cs in g["param_checksums_after_step"].items():
        if k in grads:
            gk = grads[k]
            if float(checksum(gk)[1]) < 1e-6:
                continue            # ill-conditioned: Adam amplifies pure-noise gradients
            p, _, _ = O.adam_step(sd[k], gk, torch.zeros_like(gk), torch.zeros_like(gk), 1)
        else:
            p = new_stats[k]
        mine = checksum(p)
        assert float(mine[1]) == pytest.approx(float(cs[1]), rel=2e-5, abs=1e-6), k


@pytest.mark.parametrize("name", ["diffusion", "conditional_diffusion"])
def test_sampler_trajectory(golden, name):
    """The full 1000-step loop is checked at the stored checkpoints: given the stored x_t the
    oracle must reproduce the stored eps, and one p_sample step must be bit-exact arithmetic."""
    g = golden(name)
    spec = SPECS[name]
    sd = init_state_dict(name)
    betas, alphas, ac = O.make_schedule()
    s = g["sample"]
    n = s["n"]
    cond = make_inputs(name, 4).get("cond")
    cond = None if cond is None else cond[:n]
    for t, (x, eps_ref) in s["kept"].items():
        with torch.no_grad():
            eps = O.unet_forward(spec, sd, x, torch.full((n,), t, dtype=torch.long), cond)
        assert rel(eps, eps_ref) < 1e-5, t
    # final step: x_0 = p_sample(x_1 ...) at t=0 uses no noise
    x, eps_ref = s["kept"][0]
    x0 = O.p_sample_step(x, eps_ref, None, 0, betas, alphas, ac)
    assert torch.equal(x0, s["x_0"])


def test_dit(golden):
    g = golden("diffusion_transformer")
    sd = init_state_dict("diffusion_transformer", perturb=False)
    B = g["x_t"].shape[0]
    inp = make_inputs("diffusion_transformer", B)
    _, _, ac = O.make_schedule()
    x_t = O.q_sample(ac, inp["x0"], inp["t"], inp["noise"])
    assert torch.equal(x_t, g["x_t"])
    with torch.no_grad():
        eps = O.dit_forward(sd, x_t, inp["t"], inp["cond"])
    assert rel(eps, g["eps_eval"]) < 1e-5


def test_latent_mlp_and_vae(golden):
    """latent_diffusion.NoiseModel (latent_diffusion.py:16-128) and the VAE edges (vae.py:51-62)."""
    g = golden("latent_diffusion")
    sd = init_state_dict("latent_diffusion")
    B = g["x_t"].shape[0]
    inp = make_inputs("latent_diffusion", B)
    _, _, ac = O.make_schedule()
    x_t = O.q_sample(ac, inp["x0"], inp["t"], inp["noise"])
    assert torch.equal(x_t, g["x_t"])
    with torch.no_grad():
        assert rel(O.mlp_forward(sd, x_t, inp["t"], inp["cond"]), g["eps_eval"]) < 1e-6
    fwd = lambda leaf, x, t, c, ns: O.mlp_forward(leaf, x, t, c, training=True, new_stats=ns)
    loss, grads, stats, pred = O.dense_loss_and_grads(fwd, sd, inp["x0"], inp["t"], inp["noise"], ac, inp["cond"])
    assert rel(pred, g["eps_train"]) < 1e-6
    assert float(loss) == pytest.approx(float(g["loss"]), rel=1e-6)
    for k, cs in g["grad_checksums"].items():
        if float(cs[1]) < 1e-6:
            assert float(checksum(grads[k])[1]) < 1e-5, k
        else:
            assert float(checksum(grads[k])[1]) == pytest.approx(float(cs[1]), rel=1e-4), k
    for k, v in g["buffers_after"].items():
        assert rel(stats[k].float(), v.float()) < 1e-6, k
    vsd = init_state_dict("vae", perturb=False)
    mu, logvar = O.vae_encode(vsd, g["vae"]["x"])
    assert rel(mu, g["vae"]["mu"]) < 1e-6 and rel(logvar, g["vae"]["logvar"]) < 1e-6
    assert rel(O.vae_decode(vsd, mu), g["vae"]["dec"]) < 1e-6


def test_dit_grads(golden):
    g = golden("diffusion_transformer")
    sd = init_state_dict("diffusion_transformer", perturb=False)
    inp = make_inputs("diffusion_transformer", g["x_t"].shape[0])
    _, _, ac = O.make_schedule()
    fwd = lambda leaf, x, t, c, ns: O.dit_forward(leaf, x, t, c)
    loss, grads, _, _ = O.dense_loss_and_grads(fwd, sd, inp["x0"], inp["t"], inp["noise"], ac, inp["cond"])
    assert float(loss) == pytest.approx(float(g["loss"]), rel=1e-5)
    for k, cs in g["grad_checksums"].items():
        if float(cs[1]) < 1e-9:           # Q/K projections are dead at L == 1
            assert float(checksum(grads[k])[1]) < 1e-9, k
        else:
            assert float(checksum(grads[k])[1]) == pytest.approx(float(cs[1]), rel=2e-4), k
