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


# ------------------------------------------------------------------------------------------
# classifier-free guidance (extension): no reference behaviour exists; the target is the
# composition eps_u + w*(eps_c - eps_u) of two reference (oracle) forwards inside the reference's reverse loop
# ------------------------------------------------------------------------------------------
def _cfg_model(dev, precision):
    from tinydiff.conditional_diffusion import NoiseModel
    sd10 = init_state_dict("conditional_diffusion")
    g = torch.Generator().manual_seed(5)
    sd = dict(sd10)
    sd["class_embedding.weight"] = torch.cat([sd10["class_embedding.weight"], torch.randn(1, 256, generator=g)], 0)
    model = NoiseModel(num_classes=11)                    # the 11th row is the null label
    model.load_state_dict(sd, strict=True)
    model.precision = precision
    return model.to(dev).eval(), sd


def test_cfg_step_kernel_bit_exact(dev):
    """td_psample_step_cfg against the torch composition (separate fp32 roundings): bit-exact, both halves."""
    from tinydiff import _lib as L
    from tinydiff.conditional_diffusion import ForwardProcess
    lib = L.load()
    fp = ForwardProcess()
    tab = fp._tables(dev)
    g = torch.Generator().manual_seed(3)
    n = 6 * 784
    for t, w in ((999, 3.0), (500, 1.0), (1, 7.5), (0, 2.0)):
        x = torch.randn(n, generator=g)
        eps = torch.randn(2 * n, generator=g)
        z = torch.randn(n, generator=g)
        xd = torch.cat([x, x]).to(dev)
        ed, zd = eps.to(dev), z.to(dev)
        t_dev = torch.tensor([t], dtype=torch.int32, device=dev)
        L.check(lib.td_psample_step_cfg(xd.data_ptr(), ed.data_ptr(), n, w, zd.data_ptr(), 0, tab["coef"].data_ptr(),
                                        t_dev.data_ptr(), fp.num_timesteps, None, L.stream_ptr()), "td_psample_step_cfg")
        e = eps[n:] + w * (eps[:n] - eps[n:])
        want = O.p_sample_step(x, e, z, t, fp.betas, fp.alphas, fp.alphas_cumprod)
        assert torch.equal(xd[:n].cpu(), want) and torch.equal(xd[n:].cpu(), want), (t, w)


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 3e-2)])
def test_cfg_sampler_vs_oracle_composition(dev, precision, tol):
    from tinydiff.conditional_diffusion import ForwardProcess, sample, sample_cfg
    model, sd = _cfg_model(dev, precision)
    T, n, w = 40, 4, 2.5
    fp = ForwardProcess(num_timesteps=T)
    g = torch.Generator().manual_seed(21)
    x_T = torch.randn(n, 1, 28, 28, generator=g)
    z = torch.randn(T, n, 1, 28, 28, generator=g)
    y = torch.tensor([3, 1, 4, 9])
    null = torch.full_like(y, 10)

    def eps_fn(x, t):
        tt = torch.full((n,), t, dtype=torch.long)
        ec = O.unet_forward(O.UNET_COND, sd, x, tt, y)
        eu = O.unet_forward(O.UNET_COND, sd, x, tt, null)
        return eu + w * (ec - eu)
    want, _ = O.sample_loop(eps_fn, x_T, z, fp.betas, fp.alphas, fp.alphas_cumprod)
    got = sample_cfg(model, fp, dev, n_samples=n, y=y, guidance_scale=w, x_T=x_T, z=z.to(dev))
    assert got.shape == (n, 1, 28, 28) and not model.training
    assert rel(got, want) < tol
    eager = sample_cfg(model, fp, dev, n_samples=n, y=y, guidance_scale=w, x_T=x_T, z=z.to(dev), use_graph=False)
    assert torch.equal(got, eager)
    # w = 1 is plain conditional sampling (same kernels on the conditional half, so only the combine's rounding differs)
    plain = sample(model, fp, dev, n_samples=n, y=y, x_T=x_T, z=z.to(dev))
    one = sample_cfg(model, fp, dev, n_samples=n, y=y, guidance_scale=1.0, x_T=x_T, z=z.to(dev))
    assert rel(one, plain) < (1e-4 if precision == "fp32" else 2e-2)
    with pytest.raises(ValueError):
        sample_cfg(model, fp, dev, n_samples=n, y=None)
    with pytest.raises(ValueError):
        sample_cfg(model, fp, dev, n_samples=n, y=y, null_label=11)


@pytest.mark.parametrize("precision,final_tol", [("fp32", 1e-3), ("bf16", 2e-2)])
def test_laion_sampler_vs_oracle(dev, precision, final_tol):
    """The reverse loop of conditional_diffusion_laion.py:575-587 (text-conditioned, 4x32x32 latents, sinusoidal time
    embedding, no skip resize) over all 1000 steps with x_T and the per-step noise injected, against the oracle's loop;
    `vae=None` returns the latents (the AutoencoderKL edge is out of scope)."""
    name = "conditional_diffusion_laion"
    mod, model = build(name, dev, precision)
    sd = init_state_dict(name)
    n, T = 2, 1000
    gen = torch.Generator().manual_seed(21)
    text = torch.randn(n, 768, generator=gen)
    x_T = torch.randn(n, 4, 32, 32, generator=gen)
    z = torch.randn(T, n, 4, 32, 32, generator=gen)
    fp = mod.ForwardProcess()
    eps_fn = lambda x, t: O.unet_forward(O.UNET_LAION, sd, x, torch.full((n,), t, dtype=torch.long), text)
    want, _ = O.sample_loop(eps_fn, x_T, z, fp.betas, fp.alphas, fp.alphas_cumprod)
    got = mod.sample(model, fp, dev, text_embeds=text.to(dev), x_T=x_T, z=z.to(dev))
    assert got.shape == (n, 4, 32, 32) and not model.training
    assert rel(got, want) < final_tol
    eager = mod.sample(model, fp, dev, text_embeds=text.to(dev), x_T=x_T, z=z.to(dev), use_graph=False)
    assert torch.equal(got, eager)
    with pytest.raises(ValueError):
        mod.sample(model, fp, dev)                                    # conditional_diffusion_laion.py:565-566


def test_sampler_chains_match_single_chain(dev, monkeypatch):
    """process.SamplerChains: one sample() call split into two independent sub-batches on parallel graph branches gives the
    samples of the single-chain run when the noise is injected (every sample is independent in eval mode)."""
    name = "conditional_diffusion"
    mod, model = build(name, dev, "fp32")
    fp = mod.ForwardProcess(num_timesteps=25)                 # one 20-step graph + 5 single steps per chain
    n = 64
    g = torch.Generator().manual_seed(4)
    x_T = torch.randn(n, 1, 28, 28, generator=g)
    y = torch.randint(0, 10, (n,), generator=g)
    z = torch.randn(25, n, 1, 28, 28, generator=g).to(dev)
    monkeypatch.setenv("TD_SAMPLE_CHAINS", "1")
    one = mod.sample(model, fp, dev, n_samples=n, y=y, x_T=x_T, z=z)
    monkeypatch.setenv("TD_SAMPLE_CHAINS", "2")
    two = mod.sample(model, fp, dev, n_samples=n, y=y, x_T=x_T, z=z)
    eager = mod.sample(model, fp, dev, n_samples=n, y=y, x_T=x_T, z=z, use_graph=False)
    assert rel(two, one) < 1e-5 and torch.equal(two, eager)
    a = mod.sample(model, fp, dev, n_samples=n, y=y, x_T=x_T, seed=3)       # in-kernel Philox noise, keyed per chain
    b = mod.sample(model, fp, dev, n_samples=n, y=y, x_T=x_T, seed=3)
    assert torch.isfinite(a).all() and torch.equal(a, b)
    assert not torch.equal(a[:32], a[32:])
