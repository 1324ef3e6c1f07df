"""Pin the oracle against the UNMODIFIED reference modules (build container only: the GPU box
has no /root/reference, there these tests skip and tests/test_oracle_golden.py carries the pin)."""
import pytest
import torch

from oracle import ddpm_oracle as O
from oracle import reference_shim as shim
from oracle.fixtures import (build_reference_model, init_state_dict, injected_randn, make_inputs,
                             perturb_bn)

pytestmark = pytest.mark.skipif(not shim.available(), reason="reference not mounted")

SPECS = {"diffusion": O.UNET_MNIST, "conditional_diffusion": O.UNET_COND,
         "conditional_diffusion_laion": O.UNET_LAION}


@pytest.mark.parametrize("name", list(SPECS) + ["diffusion_transformer"])
def test_seeded_init_matches_reference(name):
    ref = shim.load(name)
    model = build_reference_model(ref, name)
    sd = init_state_dict(name, perturb=False)
    rsd = model.state_dict()
    assert list(sd.keys()) == list(rsd.keys())
    for k in sd:
        assert sd[k].dtype == rsd[k].dtype and sd[k].shape == rsd[k].shape, k
        assert torch.equal(sd[k], rsd[k]), k


@pytest.mark.parametrize("name", list(SPECS))
@pytest.mark.parametrize("training", [False, True])
def test_unet_forward_bitwise(name, training):
    ref = shim.load(name)
    model = perturb_bn(build_reference_model(ref, name))
    sd = init_state_dict(name)
    inp = make_inputs(name, 3)
    model.train(training)
    args = [inp["x0"], inp["t"]] + ([inp["cond"]] if "cond" in inp else [])
    with torch.no_grad():
        want = model(*args)
        new_stats = {}
        got = O.unet_forward(SPECS[name], sd, inp["x0"], inp["t"], inp.get("cond"), training=training,
                             new_stats=new_stats)
    assert torch.equal(got, want)
    if training:
        rsd = model.state_dict()
        for k, v in new_stats.items():
            assert torch.equal(v, rsd[k]), k


def test_q_sample_and_p_sample_bitwise():
    ref = shim.load("diffusion")
    fp = ref.ForwardProcess()
    betas, alphas, ac = O.make_schedule()
    inp = make_inputs("diffusion", 5)
    with injected_randn(ref, [inp["noise"]]):
        want, _ = fp.q_sample(torch.device("cpu"), inp["x0"], inp["t"])
    assert torch.equal(O.q_sample(ac, inp["x0"], inp["t"], inp["noise"]), want)

    class Fixed(torch.nn.Module):
        def forward(self, x, t):
            return torch.sin(x * 3.0) + t.view(-1, 1, 1, 1).float() * 1e-3
    g = torch.Generator().manual_seed(5)
    short = ref.ForwardProcess(num_timesteps=7)
    x_T = torch.randn(2, 1, 28, 28, generator=g)
    zs = [torch.randn(2, 1, 28, 28, generator=g) for _ in range(6)]
    with injected_randn(ref, [x_T] + zs):
        want = ref.sample(Fixed(), short, torch.device("cpu"), n_samples=2)
    b7, a7, ac7 = O.make_schedule(7)
    z = {6 - i: zs[i] for i in range(6)}
    z[0] = None
    f = Fixed()
    got, _ = O.sample_loop(lambda x, t: f(x, torch.full((2,), t)), x_T, z, b7, a7, ac7)
    assert torch.equal(got, want)


def test_dit_matches_reference():
    ref = shim.load("diffusion_transformer")
    model = build_reference_model(ref, "diffusion_transformer").eval()
    sd = init_state_dict("diffusion_transformer", perturb=False)
    inp = make_inputs("diffusion_transformer", 16)
    with torch.no_grad():
        want = model(inp["x0"], inp["t"], inp["cond"])
        got = O.dit_forward(sd, inp["x0"], inp["t"], inp["cond"])
    assert float((got - want).norm() / want.norm()) < 2e-6


def test_adam_matches_torch_optim():
    g = torch.Generator().manual_seed(3)
    p = torch.randn(1000, generator=g)
    ref_p = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([ref_p], lr=1e-3)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    for step in range(1, 4):
        grad = torch.randn(1000, generator=g)
        ref_p.grad = grad.clone()
        opt.step()
        p, m, v = O.adam_step(p, grad, m, v, step)
        assert float((p - ref_p.detach()).abs().max()) < 1e-7
