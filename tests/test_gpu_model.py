"""Model-level parity on a real B200: NoiseModel eval forward and the reverse loop against the CPU
oracle and the reference-generated golden fixtures.  Tolerances are north_star's:
per-step eps rel-L2 <= 1e-4 (fp32 path) / <= 1e-2 (bf16 path)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ddpm_oracle as O                     # noqa: E402  (checker only)
from oracle.fixtures import init_state_dict, make_inputs  # noqa: E402

TOL = {"fp32": 1e-4, "bf16": 1e-2}


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    from tinydiff import _lib as L
    return L.require_device("cuda:0")


def build(name, dev, precision):
    import importlib
    mod = importlib.import_module(f"tinydiff.{name}")
    torch.manual_seed(0)
    model = mod.NoiseModel()
    model.load_state_dict(init_state_dict(name), strict=True)      # seeded init + perturbed BN
    model.precision = precision
    return mod, model.to(dev).eval()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["diffusion", "conditional_diffusion"])
def test_eval_forward_vs_golden(dev, golden, name, precision):
    g = golden(name)
    mod, model = build(name, dev, precision)
    inp = make_inputs(name, g["x_t"].shape[0])
    args = [g["x_t"].to(dev), inp["t"].to(dev)] + ([inp["cond"].to(dev)] if "cond" in inp else [])
    with torch.no_grad():
        eps = model(*args)
    assert eps.shape == g["eps_eval"].shape and eps.dtype == torch.float32
    assert rel(eps, g["eps_eval"]) < TOL[precision]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_eval_forward_vs_oracle_ragged_batch(dev, precision):
    """Batch 37 (not a multiple of any tile), every timestep range, against the oracle."""
    name = "conditional_diffusion"
    mod, model = build(name, dev, precision)
    sd = init_state_dict(name)
    inp = make_inputs(name, 37, seed=99)
    _, _, ac = O.make_schedule()
    x_t = O.q_sample(ac, inp["x0"], inp["t"], inp["noise"])
    want = O.unet_forward(O.UNET_COND, sd, x_t, inp["t"], inp["cond"])
    with torch.no_grad():
        got = model(x_t.to(dev), inp["t"].to(dev), inp["cond"].to(dev))
    assert rel(got, want) < TOL[precision]
    # per-sample too: no sample may be badly off
    per = ((got.cpu() - want).flatten(1).norm(dim=1) / want.flatten(1).norm(dim=1)).max()
    assert float(per) < 3 * TOL[precision]


@pytest.mark.parametrize("precision,final_tol", [("fp32", 1e-3), ("bf16", 2e-2)])
@pytest.mark.parametrize("name", ["diffusion", "conditional_diffusion"])
def test_sampler_vs_golden(dev, golden, name, precision, final_tol):
    """Full 1000-step reverse loop with the reference's x_T and per-step noise injected; compared
    with the reference's own trajectory (tests/golden).  Stated tolerance on the final sample:
    rel-L2 <= 1e-3 (fp32) / 2e-2 (bf16)."""
    g = golden(name)
    s = g["sample"]
    n, T = s["n"], 1000
    mod, model = build(name, dev, precision)
    gen = torch.Generator().manual_seed(s["seed"])
    shape = (n, 1, 28, 28)
    x_T = torch.randn(shape, generator=gen)
    zs = [torch.randn(shape, generator=gen) for _ in range(T - 1)]
    z = torch.zeros((T,) + shape)
    for i, zz in enumerate(zs):
        z[T - 1 - i] = zz                         # the i-th draw is used at t = T-1-i
    fp = mod.ForwardProcess()
    kw = {}
    if name == "conditional_diffusion":
        kw["y"] = make_inputs(name, 4)["cond"][:n].to(dev)
    x0 = mod.sample(model, fp, dev, n_samples=n, x_T=x_T, z=z.to(dev), **kw)
    assert not model.training
    assert rel(x0, s["x_0"]) < final_tol
    # eager (no CUDA graph) path gives the same bits
    x0b = mod.sample(model, fp, dev, n_samples=n, x_T=x_T, z=z.to(dev), use_graph=False, **kw)
    assert torch.equal(x0, x0b)


def test_sampler_philox_runs_and_is_seeded(dev):
    mod, model = build("diffusion", dev, "bf16")
    fp = mod.ForwardProcess(num_timesteps=50)
    x_T = torch.randn(4, 1, 28, 28, generator=torch.Generator().manual_seed(1))
    a = mod.sample(model, fp, dev, n_samples=4, x_T=x_T, seed=11)
    b = mod.sample(model, fp, dev, n_samples=4, x_T=x_T, seed=11)
    c = mod.sample(model, fp, dev, n_samples=4, x_T=x_T, seed=12)
    assert torch.isfinite(a).all() and torch.equal(a, b) and not torch.equal(a, c)


def test_conditional_sample_errors(dev):
    mod, model = build("conditional_diffusion", dev, "bf16")
    fp = mod.ForwardProcess()
    with pytest.raises(ValueError):
        mod.sample(model, fp, dev, n_samples=4)                       # conditional_diffusion.py:358-361
    with pytest.raises(ValueError):
        mod.sample(model, fp, dev, n_samples=4, y=torch.zeros(3, dtype=torch.long))   # :362-363
