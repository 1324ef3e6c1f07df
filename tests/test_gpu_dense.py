"""Parity of the two latent denoisers on a real B200: the MLP "U-Net" (latent_diffusion.py:16-128)
and the length-1 "DiT" (diffusion_transformer.py:16-109), plus the VAE edges (vae.py:51-62).
All fp32: eps / loss / gradients within 1e-4 of the CPU oracle (5e-4 for the BatchNorm1d MLP whose
tiny-batch statistics amplify rounding)."""
import importlib

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import ddpm_oracle as O                        # noqa: E402  (checker only)
from oracle.fixtures import init_state_dict, make_inputs   # noqa: E402


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a B200"
    from tinydiff import _lib as L
    return L.require_device("cuda:0")


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def build(name, dev, train=False, **kw):
    mod = importlib.import_module(f"tinydiff.{name}")
    model = mod.NoiseModel(**kw)
    sd = init_state_dict(name, perturb=(name == "latent_diffusion"))
    model.load_state_dict(sd, strict=True)
    model = model.to(dev)
    return mod, (model.train() if train else model.eval()), sd


FWD = {"latent_diffusion": lambda leaf, x, t, c, ns, tr=False: O.mlp_forward(leaf, x, t, c, training=tr, new_stats=ns),
       "diffusion_transformer": lambda leaf, x, t, c, ns, tr=False: O.dit_forward(leaf, x, t, c)}
KW = {"latent_diffusion": {}, "diffusion_transformer": {"dropout": 0.0}}


@pytest.mark.parametrize("name", ["latent_diffusion", "diffusion_transformer"])
def test_eval_forward_vs_golden(dev, golden, name):
    g = golden(name)
    mod, model, sd = build(name, dev, **KW[name])
    inp = make_inputs(name, g["x_t"].shape[0])
    with torch.no_grad():
        eps = model(g["x_t"].to(dev), inp["t"].to(dev), inp["cond"].to(dev))
    assert eps.shape == g["eps_eval"].shape
    assert rel(eps, g["eps_eval"]) < 1e-5
    # ragged batch against the oracle
    inp = make_inputs(name, 37, seed=5)
    with torch.no_grad():
        got = model(inp["x0"].to(dev), inp["t"].to(dev), inp["cond"].to(dev))
        want = FWD[name](sd, inp["x0"], inp["t"], inp["cond"], None)
    assert rel(got, want) < 1e-5


@pytest.mark.parametrize("name", ["latent_diffusion", "diffusion_transformer"])
def test_train_step_autograd_vs_oracle(dev, golden, name):
    """q_sample -> model -> mse_loss -> backward -> Adam with the drop-in classes
    (latent_diffusion.py:207-223 / diffusion_transformer.py:189-202)."""
    g = golden(name)
    mod, model, sd = build(name, dev, train=True, **KW[name])
    B = g["x_t"].shape[0]
    inp = make_inputs(name, B)
    fp = mod.ForwardProcess()
    fwd = lambda leaf, x, t, c, ns: FWD[name](leaf, x, t, c, ns, True)
    loss_ref, grads_ref, stats_ref, pred_ref = O.dense_loss_and_grads(fwd, sd, inp["x0"], inp["t"], inp["noise"],
                                                                     fp.alphas_cumprod, inp["cond"])
    opt = torch.optim.Adam(model.parameters(), lr=3e-4)
    x_t, noise = fp.q_sample(dev, inp["x0"], inp["t"].to(dev), noise=inp["noise"])
    assert torch.equal(x_t.cpu(), g["x_t"])
    pred = model(x_t, inp["t"].to(dev), inp["cond"].to(dev))
    loss = F.mse_loss(pred, noise)
    opt.zero_grad()
    loss.backward()
    tol = 5e-4 if name == "latent_diffusion" else 1e-4
    assert rel(pred, pred_ref) < tol
    assert abs(float(loss) - float(g["loss"])) / float(g["loss"]) < 1e-5
    bad = {}
    for k, p in model.named_parameters():
        ref = grads_ref[k]
        if float(ref.norm()) < 1e-7:
            # Linear biases in front of a train-mode BatchNorm1d (MLP); everything else must be exactly zero
            assert float(p.grad.abs().max()) < 1e-6, k
            continue
        if k.endswith("in_proj_weight") or k.endswith("in_proj_bias"):
            D = 256
            assert float(p.grad[:2 * D].abs().max()) == 0.0, k        # dead Q / K projections
        e = rel(p.grad, ref)
        if e > 2e-3:
            bad[k] = e
    assert not bad, bad
    for k, v in stats_ref.items():
        got = dict(model.named_buffers())[k]
        assert rel(got.float(), v.float()) < 1e-5, k
    opt.step()


@pytest.mark.parametrize("name", ["latent_diffusion", "diffusion_transformer"])
@pytest.mark.parametrize("B,passes", [(128, "1"), (8, "1"), (135, "1"), (1, "1"), (200, "0"), (251, "0")])
def test_cluster_kernel_vs_oracle(dev, name, B, passes, monkeypatch):
    """The cluster-resident eval forward (csrc/dense_cluster.cu) against the CPU oracle at the benchmark batch (9 rows per cluster
    on 15 clusters), at ragged / tiny batches, and -- forced, the engine would pick the tape there -- at batches that take a
    second pass over the weights; the grid-barrier tape and the one-launch-per-op path must agree with it to rounding."""
    from tinydiff.dense import DenseEngine
    mod, model, sd = build(name, dev, **KW[name])
    inp = make_inputs(name, B, seed=11)
    want = FWD[name](sd, inp["x0"], inp["t"], inp["cond"], None)
    outs = {}
    for mode in ("1", "2", "0"):
        monkeypatch.setenv("TD_DENSE_FUSED", mode)
        monkeypatch.setenv("TD_DENSE_CLUSTER_PASSES", passes)
        e = DenseEngine(model, B, dev, False, model.in_dim, model.emb_mode)
        model._declare(e)
        e.build()
        assert (e._ctapes is not None) == (mode == "1"), "the cluster kernel must be the path that runs"
        e.load_inputs(inp["x0"].to(dev), inp["t"].to(dev), inp["cond"].to(dev))
        for _ in range(2):                       # twice: the mbarrier phases / weight ring must be reusable across launches
            e.launch_forward()
        outs[mode] = e.eps.clone()
    assert rel(outs["1"], want) < 1e-5
    assert rel(outs["1"], outs["2"]) < 2e-6 and rel(outs["1"], outs["0"]) < 2e-6


@pytest.mark.parametrize("name", ["latent_diffusion", "diffusion_transformer"])
def test_cluster_merged_weights_follow_parameter_updates(dev, name):
    """Back-to-back Linears are merged into one in the cluster tape (W = W_b W_a, derived buffers): an in-place parameter update
    must reach the next forward and the next sampler call."""
    mod, model, sd = build(name, dev, **KW[name])
    eng = model.engine(16, dev, training=False)
    assert eng._ctapes is not None and len(eng._merged) >= 1
    inp = make_inputs(name, 16, seed=3)
    args = (inp["x0"].to(dev), inp["t"].to(dev), inp["cond"].to(dev))
    with torch.no_grad():
        eps0 = model(*args)
        for mg in eng._merged:                       # touch both factors of every merged pair
            mg["wb"].mul_(1.25)
            if mg["ba"] is not None:
                mg["ba"].add_(0.05)
        eps1 = model(*args)
    sd2 = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    want = FWD[name](sd2, inp["x0"], inp["t"], inp["cond"], None)
    assert rel(eps1, want) < 1e-5
    assert rel(eps0, eps1) > 1e-4, "the update must change the output"
    # TrainStep-style updates go through raw pointers: the model-level generation counter invalidates the merged weights too
    with torch.no_grad():
        for mg in eng._merged:
            mg["wa"].data.mul_(0.9)                  # .data: no _version bump, like a kernel writing through the pointer
    model._weights_gen = getattr(model, "_weights_gen", 0) + 1
    with torch.no_grad():
        eps2 = model(*args)
    sd3 = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    assert rel(eps2, FWD[name](sd3, inp["x0"], inp["t"], inp["cond"], None)) < 1e-5


@pytest.mark.parametrize("name", ["latent_diffusion", "diffusion_transformer"])
@pytest.mark.parametrize("n", [37, 128])
def test_fused_reverse_step_matches_unfused(dev, name, n, monkeypatch):
    """The reverse-step update folded into the cluster kernel's last epilogue (td_dense_cluster_step: one launch per step) is
    BIT-identical to the tape followed by td_psample_step_advance -- injected noise table and in-kernel Philox noise, eager and
    captured -- and leaves the step counter at -1."""
    from tinydiff import _lib as L
    from tinydiff.dense import dense_sample
    mod, model, sd = build(name, dev, **KW[name])
    T = 12
    fp = mod.ForwardProcess(num_timesteps=T)
    g = torch.Generator().manual_seed(n)
    x_T = torch.randn(n, 20, generator=g)
    z = torch.randn(T, n, 20, generator=g).to(dev)
    y = torch.randint(0, 10, (n,), generator=g).to(dev)
    lib = L.load()
    outs = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("TD_DENSE_STEP_FUSED", fused)
        model._init_engines()
        for use_graph in (False, True):
            outs[(fused, "z", use_graph)] = dense_sample(None, model, fp, dev, n, y, x_T=x_T, z=z, use_graph=use_graph)
            outs[(fused, "seed", use_graph)] = dense_sample(None, model, fp, dev, n, y, x_T=x_T, seed=77, use_graph=use_graph)
        eng = model.engine(n, dev, training=False)
        assert int(eng.t_dev.item()) == -1
        assert eng._reverse_loop.launches_per_step == (1 if fused == "1" else 2)
    for kind in ("z", "seed"):
        for use_graph in (False, True):
            assert torch.equal(outs[("1", kind, use_graph)], outs[("0", kind, use_graph)]), (kind, use_graph)
    assert not torch.equal(outs[("1", "z", True)], outs[("1", "seed", True)])


def test_cluster_kernel_is_default_at_reference_batch(dev):
    """At the reference batch the public forward runs the cluster kernel (one launch), above one pass the tape."""
    from tinydiff import _lib as L
    mod, model, sd = build("diffusion_transformer", dev, dropout=0.0)
    lib = L.load()
    for B, want_cluster in ((128, True), (200, False)):
        inp = make_inputs("diffusion_transformer", B)
        with torch.no_grad():
            model(inp["x0"].to(dev), inp["t"].to(dev), inp["cond"].to(dev))
            n0 = int(lib.td_launch_count())
            model(inp["x0"].to(dev), inp["t"].to(dev), inp["cond"].to(dev))
        assert int(lib.td_launch_count()) - n0 == 1
        eng = model.engine(B, dev, training=False)
        assert (eng._ctapes is not None) == want_cluster and (eng._tapes is not None) == (not want_cluster)


@pytest.mark.parametrize("name", ["latent_diffusion", "diffusion_transformer"])
def test_sampler_vs_oracle(dev, name):
    """Reverse loop with injected x_T / z (short schedule) against the oracle's sample_loop."""
    mod, model, sd = build(name, dev, **KW[name])
    T, n = 50, 5
    fp = mod.ForwardProcess(num_timesteps=T)
    betas, alphas, ac = O.make_schedule(T)
    g = torch.Generator().manual_seed(3)
    x_T = torch.randn(n, 20, generator=g)
    z = torch.randn(T, n, 20, generator=g)
    y = torch.randint(0, 10, (n,), generator=g)
    zs = [None] + [z[t] for t in range(1, T)]
    want, _ = O.sample_loop(lambda x, t: FWD[name](sd, x, torch.full((n,), t), y, None), x_T, zs, betas, alphas, ac)
    for use_graph in (False, True):
        from tinydiff.dense import dense_sample
        got = dense_sample(None, model, fp, dev, n, y.to(dev), x_T=x_T, z=z.to(dev), use_graph=use_graph)
        assert rel(got, want) < 1e-4, use_graph
    with pytest.raises(ValueError):
        mod.sample(None, model, fp, dev, n_samples=n, y=None)
    with pytest.raises(ValueError):
        mod.sample(None, model, fp, dev, n_samples=n + 1, y=y.to(dev))


@pytest.mark.parametrize("name", ["latent_diffusion", "diffusion_transformer"])
@pytest.mark.parametrize("use_graph", [False, True])
def test_fused_train_step(dev, name, use_graph):
    from tinydiff.train import TrainStep
    mod, model, sd0 = build(name, dev, train=True, **KW[name])
    B = 16
    fp = mod.ForwardProcess()
    ts = TrainStep(model, fp, B, dev, lr=3e-4, use_graph=use_graph)
    sd = {k: v.clone() for k, v in sd0.items()}
    m = {k: torch.zeros_like(v) for k, v in sd.items() if O.is_param(k)}
    v_ = {k: torch.zeros_like(v) for k, v in sd.items() if O.is_param(k)}
    fwd = lambda leaf, x, t, c, ns: FWD[name](leaf, x, t, c, ns, True)
    for step in (1, 2):
        inp = make_inputs(name, B, seed=700 + step)
        loss_ref, grads, stats, _ = O.dense_loss_and_grads(fwd, sd, inp["x0"], inp["t"], inp["noise"], fp.alphas_cumprod,
                                                          inp["cond"])
        loss = ts(inp["x0"], inp["cond"], t=inp["t"], noise=inp["noise"])
        assert abs(float(loss) - float(loss_ref)) / float(loss_ref) < 2e-4
        for k in m:
            g_ = grads[k]
            if float(g_.norm()) < 1e-7:
                g_ = torch.zeros_like(g_)
            sd[k], m[k], v_[k] = O.adam_step(sd[k], g_, m[k], v_[k], step, lr=3e-4)
        sd.update(stats)
    got = model.state_dict()
    for k in m:
        upd_ref = sd[k] - sd0[k]
        if float(upd_ref.norm()) == 0:
            continue
        assert rel(got[k].cpu() - sd0[k], upd_ref) < 5e-2, k
        assert rel(got[k], sd[k]) < 1e-4, k


def test_dit_dropout_train_mode(dev):
    """dropout = 0.05 (the reference default) in train mode: masks are Bernoulli(0.95)/0.95, shared per
    (sample, head) on the attention branch; eval mode is unaffected; gradients stay finite and the
    dead Q/K rows stay exactly zero."""
    import ctypes as C
    from tinydiff import _lib as L
    lib = L.load()
    rows, cols, group, p = 512, 256, 64, 0.05
    x = torch.ones(rows, cols, device=dev)
    out = torch.empty_like(x)
    seed = torch.tensor([1234, 7], dtype=torch.int64, device=dev)
    L.check(lib.td_dropout_f32(x.data_ptr(), cols, out.data_ptr(), cols, rows, cols, group, p, seed.data_ptr(),
                               L.stream_ptr()))
    vals = out.unique()
    assert set(round(float(v), 5) for v in vals) <= {0.0, round(1 / 0.95, 5)}
    assert torch.equal(out.view(rows, cols // group, group).amax(-1), out.view(rows, cols // group, group).amin(-1))
    keep = float((out > 0).float().mean())
    assert abs(keep - 0.95) < 0.02
    out2 = torch.empty_like(x)
    L.check(lib.td_dropout_f32(x.data_ptr(), cols, out2.data_ptr(), cols, rows, cols, group, p, seed.data_ptr(),
                               L.stream_ptr()))
    assert torch.equal(out, out2)
    mod, model, sd = build("diffusion_transformer", dev, train=True, dropout=0.05)
    inp = make_inputs("diffusion_transformer", 64)
    args = (inp["x0"].to(dev), inp["t"].to(dev), inp["cond"].to(dev))
    torch.manual_seed(1)
    a = model(*args)
    b = model(*args)
    assert rel(a, b) > 1e-3                                   # different masks per call
    loss = F.mse_loss(b, inp["noise"].to(dev))              # the latest forward owns the engine's saved tensors
    loss.backward()
    with pytest.raises(RuntimeError, match="overwritten by a later train-mode forward"):
        F.mse_loss(a, inp["noise"].to(dev)).backward()
    assert all(torch.isfinite(q.grad).all() for q in model.parameters())
    assert float(model.transformer_blocks[0].attention.in_proj_weight.grad[:512].abs().max()) == 0.0
    want = O.dit_forward(sd, inp["x0"], inp["t"], inp["cond"])
    assert 1e-3 < rel(a, want) < 0.5                          # a perturbation of the dropout-free output
    with torch.no_grad():
        assert rel(model.eval()(*args), want) < 1e-5


def test_vae_edges(dev, golden):
    from tinydiff.vae import VAE
    g = golden("latent_diffusion")["vae"]
    vae = VAE()
    vae.load_state_dict(init_state_dict("vae", perturb=False), strict=True)
    vae = vae.to(dev).eval()
    mu, logvar = vae.encode(g["x"].to(dev))
    assert rel(mu, g["mu"]) < 1e-5 and rel(logvar, g["logvar"]) < 1e-5
    assert rel(vae.decode(mu), g["dec"]) < 1e-5
    eps = torch.randn_like(mu)
    assert rel(vae.reparameterize(mu, logvar, eps), O.vae_reparameterize(mu.cpu(), logvar.cpu(), eps.cpu())) < 1e-6
    # the sampler tail: decode(z).view(-1, 1, 28, 28)   (latent_diffusion.py:346)
    from tinydiff.latent_diffusion import ForwardProcess, NoiseModel, sample
    model = NoiseModel()
    model.load_state_dict(init_state_dict("latent_diffusion"))
    model = model.to(dev)
    x = sample(vae, model, ForwardProcess(num_timesteps=5), dev, n_samples=3, y=torch.tensor([1, 2, 3], device=dev), seed=1)
    assert x.shape == (3, 1, 28, 28) and float(x.min()) >= 0 and float(x.max()) <= 1


@pytest.mark.parametrize("M,D", [(8, 256), (37, 20), (128, 1024), (8192, 256)])
def test_layernorm_kernels(dev, M, D):
    import ctypes as C
    from tinydiff import _lib as L
    lib = L.load()
    g = torch.Generator().manual_seed(M)
    x = torch.randn(M, D, generator=g) * 3 + 1
    gamma, beta = torch.rand(D, generator=g) + 0.5, torch.randn(D, generator=g)
    dy = torch.randn(M, D, generator=g)
    xr, gr, br = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = F.layer_norm(xr, (D,), gr, br)
    yr.backward(dy)
    xd, gd, bd, dyd = x.to(dev), gamma.to(dev), beta.to(dev), dy.to(dev)
    y, mean, rstd = torch.empty_like(xd), torch.empty(M, device=dev), torch.empty(M, device=dev)
    st = L.stream_ptr()
    L.check(lib.td_layernorm_fwd(xd.data_ptr(), gd.data_ptr(), bd.data_ptr(), y.data_ptr(), mean.data_ptr(),
                                 rstd.data_ptr(), M, D, 1e-5, st))
    assert rel(y, yr) < 1e-5
    dx, dg, db = torch.empty_like(xd), torch.empty(D, device=dev), torch.empty(D, device=dev)
    ws = torch.empty(2 * D * 64, device=dev) if M >= 4096 else None         # many-CTA parameter gradients at large batch
    L.check(lib.td_layernorm_bwd(dyd.data_ptr(), xd.data_ptr(), gd.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                 dx.data_ptr(), dg.data_ptr(), db.data_ptr(), M, D, L.ptr(ws), ws.numel() if ws is not None else 0, st))
    assert rel(dx, xr.grad) < 1e-5 and rel(dg, gr.grad) < 1e-5 and rel(db, br.grad) < 1e-5


@pytest.mark.parametrize("M,rows,use_ws", [(100, 10, False), (8192, 10, True), (5000, 1, True)])
def test_embedding_bwd(dev, M, rows, use_ws):
    """nn.Embedding backward: one CTA per table row, and the many-CTA chunked form (workspace) at large batch."""
    from tinydiff import _lib as L
    lib = L.load()
    D = 256
    g = torch.Generator().manual_seed(M + rows)
    grad = torch.randn(M, D, generator=g)
    idx = torch.randint(0, rows, (M,), generator=g)
    want = torch.zeros(rows, D).index_add_(0, idx, grad)
    gd, idd = grad.to(dev), idx.to(dev)
    out = torch.full((rows, D), float("nan"), device=dev)
    ws = torch.empty(64 * rows * D, device=dev) if use_ws else None
    L.check(lib.td_embedding_bwd(gd.data_ptr(), D, idd.data_ptr(), out.data_ptr(), M, D, rows, 0, L.ptr(ws),
                                 ws.numel() if ws is not None else 0, L.stream_ptr()))
    assert rel(out, want) < 1e-5


@pytest.mark.parametrize("M,N,relu", [(8, 64, True), (128, 512, True), (37, 100, False)])
def test_bn1d_kernels(dev, M, N, relu):
    from tinydiff import _lib as L
    lib = L.load()
    g = torch.Generator().manual_seed(N)
    x = torch.randn(M, N, generator=g) * 2 + 5
    gamma, beta = torch.rand(N, generator=g) + 0.5, torch.randn(N, generator=g) * 0.1
    rm, rv = torch.randn(N, generator=g), torch.rand(N, generator=g) + 0.5
    dy = torch.randn(M, N, generator=g)
    xr, gr, br = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm_r, rv_r = rm.clone(), rv.clone()
    yr = F.batch_norm(xr, rm_r, rv_r, gr, br, True, 0.1, 1e-5)
    yr = F.relu(yr) if relu else yr
    yr.backward(dy)
    d = lambda t: t.to(dev)
    xd, gd, bd, rmd, rvd, dyd = d(x), d(gamma), d(beta), d(rm), d(rv), d(dy)
    y, sm, sr = torch.empty_like(xd), torch.empty(N, device=dev), torch.empty(N, device=dev)
    st = L.stream_ptr()
    L.check(lib.td_bn1d_fwd(xd.data_ptr(), N, gd.data_ptr(), bd.data_ptr(), rmd.data_ptr(), rvd.data_ptr(), sm.data_ptr(),
                            sr.data_ptr(), y.data_ptr(), N, M, N, 1e-5, 0.1, 1, int(relu), st))
    assert rel(y, yr) < 1e-5 and rel(rmd, rm_r) < 1e-6 and rel(rvd, rv_r) < 1e-5
    dx, dg, db = torch.empty_like(xd), torch.empty(N, device=dev), torch.empty(N, device=dev)
    L.check(lib.td_bn1d_bwd(dyd.data_ptr(), N, xd.data_ptr(), N, y.data_ptr(), N, gd.data_ptr(), sm.data_ptr(),
                            sr.data_ptr(), dx.data_ptr(), N, dg.data_ptr(), db.data_ptr(), M, N, int(relu), st))
    assert rel(dx, xr.grad) < 2e-5 and rel(dg, gr.grad) < 2e-5 and rel(db, br.grad) < 1e-5
    # eval mode uses the running statistics
    ye = torch.empty_like(xd)
    L.check(lib.td_bn1d_fwd(xd.data_ptr(), N, gd.data_ptr(), bd.data_ptr(), rmd.data_ptr(), rvd.data_ptr(), None, None,
                            ye.data_ptr(), N, M, N, 1e-5, 0.1, 0, int(relu), st))
    want = F.batch_norm(x, rm_r, rv_r, gamma, beta, False, 0.1, 1e-5)
    assert rel(ye, F.relu(want) if relu else want) < 1e-5


# ------------------------------------------------------------------------------------------
# large-batch path: tcgen05 kind::tf32 GEMM (linear_tc.cu), M >= 2048
# ------------------------------------------------------------------------------------------
def _gemm_args(A, W, out, bias=None, act=0, pre=None, residual=None, gidx=None, gtab=None, accumulate=0):
    from tinydiff import _lib as L
    g = L.GemmArgs()
    M, K = A.shape
    N = W.shape[0]
    g.M, g.N, g.K, g.alpha = M, N, K, 1.0
    g.A, g.a_rs, g.a_cs = A.data_ptr(), A.stride(0), 1
    g.B, g.b_rs, g.b_cs = W.data_ptr(), 1, W.stride(0)
    g.C, g.ldc, g.bias, g.act = out.data_ptr(), out.stride(0), L.ptr(bias), act
    g.pre_out, g.ld_pre = L.ptr(pre), (pre.stride(0) if pre is not None else 0)
    g.residual, g.ldr = L.ptr(residual), (residual.stride(0) if residual is not None else 0)
    g.gather_idx, g.gather_table, g.ld_table = L.ptr(gidx), L.ptr(gtab), (gtab.stride(0) if gtab is not None else 0)
    g.accumulate, g.splitk_ws = accumulate, None
    g.allow_tf32 = 1
    return g


@pytest.mark.parametrize("M,N,K", [(2048, 64, 64), (4096, 256, 256), (4100, 520, 1000), (2500, 36, 40), (8192, 1024, 256)])
def test_gemm_tf32_large_batch(dev, M, N, K):
    """y = act(x W^T + b) + residual + table[idx] (+ C) on the tensor-core path against fp64; ragged M / N / K tails
    (TMA zero fill), strided A (a view into a wider buffer), every epilogue option.  TF32 operands: rel-L2 <= 2e-3."""
    import ctypes as C
    from tinydiff import _lib as L
    lib = L.load()
    g_ = torch.Generator().manual_seed(M + N + K)
    wide = torch.randn(M, K + 8, generator=g_).to(dev)
    A = wide[:, 4:4 + K]                                    # row stride K + 8, 16-byte aligned start
    W = (torch.randn(N, K, generator=g_) / K ** 0.5).to(dev)
    bias = torch.randn(N, generator=g_).to(dev)
    res = torch.randn(M, N, generator=g_).to(dev)
    tab = torch.randn(10, N, generator=g_).to(dev)
    idx = torch.randint(0, 10, (M,), generator=g_).to(dev)
    ref_pre = A.double() @ W.double().t() + bias.double()
    for act, name in ((L.ACT_NONE, "none"), (L.ACT_GELU, "gelu"), (L.ACT_SILU, "silu"), (L.ACT_RELU, "relu")):
        out = torch.full((M, N), 0.5, device=dev)
        pre = torch.empty(M, N, device=dev)
        g = _gemm_args(A, W, out, bias=bias, act=act, pre=pre, residual=res, gidx=idx, gtab=tab, accumulate=1)
        assert lib.td_gemm_f32_path(C.byref(g)) == 1, "not on the tensor-core path"
        L.check(lib.td_gemm_f32(C.byref(g), L.stream_ptr()), "td_gemm_f32")
        a = {"none": lambda v: v, "gelu": lambda v: F.gelu(v), "silu": F.silu, "relu": F.relu}[name](ref_pre)
        want = a + res.double() + tab.double()[idx] + 0.5
        assert rel(pre, ref_pre) < 2e-3, name
        assert rel(out, want) < 2e-3, name
    # plain product, no epilogue; and the small-batch call stays on the exact fp32 kernels
    out = torch.empty(M, N, device=dev)
    g = _gemm_args(A, W, out)
    L.check(lib.td_gemm_f32(C.byref(g), L.stream_ptr()), "td_gemm_f32")
    assert rel(out, A.double() @ W.double().t()) < 2e-3
    exact = _gemm_args(A, W, out)
    exact.allow_tf32 = 0                                    # the default of every other caller: exact fp32 products
    assert lib.td_gemm_f32_path(C.byref(exact)) == 0
    L.check(lib.td_gemm_f32(C.byref(exact), L.stream_ptr()), "td_gemm_f32")
    assert rel(out, A.double() @ W.double().t()) < 1e-5


@pytest.mark.parametrize("name", ["latent_diffusion", "diffusion_transformer"])
def test_eval_forward_large_batch_tf32(dev, name):
    """Eval forward at 4096 samples (tensor-core GEMMs) against the CPU oracle: rel-L2 <= 1e-2, the tolerance of the other
    reduced-precision engine; the same model at 128 samples stays within the fp32 tolerance."""
    mod, model, sd = build(name, dev, **KW[name])
    B = 4096
    inp = make_inputs(name, B)
    with torch.no_grad():
        got = model(inp["noise"].to(dev), inp["t"].to(dev), inp["cond"].to(dev))
        want = FWD[name](sd, inp["noise"], inp["t"], inp["cond"], None)
        small = model(inp["noise"][:128].to(dev), inp["t"][:128].to(dev), inp["cond"][:128].to(dev))
    assert torch.isfinite(got).all()
    assert rel(got, want) < 1e-2
    assert rel(small, want[:128]) < 5e-4


@pytest.mark.parametrize("name", ["latent_diffusion", "diffusion_transformer"])
def test_train_step_large_batch_tf32(dev, name):
    """Fused train step at 2048 samples: forward and data-gradient GEMMs on the tensor-core path (transposed weight copies),
    weight gradients on the FFMA kernels.  Loss and updated parameters against the CPU oracle's Adam step: 1e-2 / update
    direction agreement (Adam's first step is lr * sign(g))."""
    from tinydiff.train import TrainStep
    mod, model, sd = build(name, dev, train=True, **KW[name])
    B = 2048
    fp = mod.ForwardProcess()
    inp = make_inputs(name, B)
    ts = TrainStep(model, fp, B, dev, lr=1e-3, use_graph=True)
    loss = float(ts(inp["x0"], inp["cond"], t=inp["t"], noise=inp["noise"]))
    fwd = (lambda leaf, x, t, c, ns: FWD[name](leaf, x, t, c, ns, True))
    want_loss, grads, _, _ = O.dense_loss_and_grads(fwd, sd, inp["x0"], inp["t"], inp["noise"], fp.alphas_cumprod, inp["cond"])
    assert abs(loss - float(want_loss)) / float(want_loss) < 1e-2
    after = model.state_dict()
    agree, total = 0, 0
    for k, g in grads.items():
        if float(g.norm()) < 1e-7:                            # Linear biases in front of a BatchNorm1d: zero up to rounding noise
            continue
        big = g.abs() > 0.1 * g.abs().max()                  # entries whose sign is not at the mercy of rounding
        step = (after[k].cpu() - sd[k])[big]
        agree += int((torch.sign(step) == -torch.sign(g[big])).sum())
        total += int(big.sum())
    assert total > 1000 and agree / total > 0.99, (agree, total)


@pytest.mark.parametrize("rows,cols", [(2048, 256), (4100, 20), (33, 1000)])
def test_transpose_kernel(dev, rows, cols):
    from tinydiff import _lib as L
    lib = L.load()
    src = torch.randn(rows, cols + 4, device=dev)[:, 2:2 + cols]            # strided source
    dst = torch.zeros(cols, rows + 8, device=dev)
    L.check(lib.td_transpose_f32(src.data_ptr(), src.stride(0), dst.data_ptr(), dst.stride(0), rows, cols, L.stream_ptr()),
            "td_transpose_f32")
    assert torch.equal(dst[:, :rows], src.t())
    assert float(dst[:, rows:].abs().max()) == 0.0


def test_weight_gradient_gemm_takes_tensor_core_path(dev):
    """dW = g^T x with the batch as the reduction dimension: eligible once both operands are transposed (K-major)."""
    import ctypes as C
    from tinydiff import _lib as L
    lib = L.load()
    B, N, K = 4096, 256, 128
    g_ = torch.Generator().manual_seed(3)
    g = torch.randn(B, N, generator=g_).to(dev)
    x = torch.randn(B, K, generator=g_).to(dev)
    gt, xt = g.t().contiguous(), x.t().contiguous()
    out = torch.empty(N, K, device=dev)
    a = _gemm_args(gt, xt, out)                      # A = g^T [N][B], "W" = x^T [K][B]: out = g^T x
    assert lib.td_gemm_f32_path(C.byref(a)) == 1
    L.check(lib.td_gemm_f32(C.byref(a), L.stream_ptr()), "td_gemm_f32")
    assert rel(out, g.double().t() @ x.double()) < 2e-3


@pytest.mark.parametrize("M,N,K", [(256, 256, 16384), (1024, 256, 65536), (64, 512, 8192)])
def test_gemm_tf32_splitk(dev, M, N, K):
    """The tcgen05 kind::tf32 GEMM with K slices (the batch-reducing weight gradients of the dense denoisers at large batch):
    C = A B^T with both operands K-major, raw partials + fixed-order second pass; tf32 tolerance, bit-reproducible."""
    import ctypes as C
    from tinydiff import _lib as L
    lib = L.load()
    g = torch.Generator().manual_seed(K)
    A = torch.randn(M, K, generator=g)
    Bm = torch.randn(N, K, generator=g)
    bias = torch.randn(N, generator=g)
    want = A.double() @ Bm.double().t() + bias.double()
    Ad, Bd, bd = A.to(dev), Bm.to(dev), bias.to(dev)
    need = int(lib.td_gemm_f32_workspace(M, N, K))
    assert need >= 2 * M * N
    ws = torch.empty(need, device=dev)
    outs = []
    for _ in range(2):
        out = torch.empty(M, N, device=dev)
        ga = L.GemmArgs()
        ga.M, ga.N, ga.K, ga.alpha = M, N, K, 1.0
        ga.A, ga.a_rs, ga.a_cs = Ad.data_ptr(), K, 1
        ga.B, ga.b_rs, ga.b_cs = Bd.data_ptr(), 1, K
        ga.C, ga.ldc, ga.bias = out.data_ptr(), N, bd.data_ptr()
        ga.splitk_ws, ga.allow_tf32 = ws.data_ptr(), 1
        assert int(lib.td_gemm_f32_path(C.byref(ga))) == 1
        L.check(lib.td_gemm_f32(C.byref(ga), L.stream_ptr()), "td_gemm_f32")
        outs.append(out)
    assert rel(outs[0], want) < 2e-3
    assert torch.equal(outs[0], outs[1])
