"""SURVEY.md row A10 on the device: the fused multi-tensor Adam kernel with injected gradients against
``torch.optim.Adam``; the global-norm clip kernel against ``torch.nn.utils.clip_grad_norm_``; and the whole optimizer
path of ``TrainStep`` (clip -> Adam, learning rate driven by the reference's own ``CosineAnnealingLR`` lines through
``TrainStep.optimizer``) replayed on the CPU with torch's optimizer on the gradients the step produced.  Also the
device-side RNG of the train step and the bounds of the reverse-step kernel."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ddpm_oracle as O                       # noqa: E402  (checker only)
from oracle.fixtures import init_state_dict, make_inputs  # noqa: E402


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    from tinydiff import _lib as L
    return L.require_device("cuda:0")


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _adam_tables(ps, gs, ms, vs, dev, CH):
    t64 = lambda v: torch.tensor(v, dtype=torch.int64, device=dev)
    ct, co = [], []
    for i, p in enumerate(ps):
        for c in range(0, p.numel(), CH):
            ct.append(i)
            co.append(c)
    return (t64([p.data_ptr() for p in ps]), t64([g.data_ptr() for g in gs]), t64([m.data_ptr() for m in ms]),
            t64([v.data_ptr() for v in vs]), t64([p.numel() for p in ps]),
            torch.tensor(ct, dtype=torch.int32, device=dev), t64(co), len(ct))


@pytest.mark.parametrize("use_lr_dev", [False, True])
def test_adam_multi_injected_gradients_vs_torch(dev, use_lr_dev):
    """td_adam_multi on ragged tensors (scalar tails, several chunks per tensor), three steps with fresh injected gradients,
    a gradient scale and a learning rate that changes between steps, against torch.optim.Adam: <= 1e-6."""
    from tinydiff import _lib as L
    lib = L.load()
    g = torch.Generator().manual_seed(7)
    shapes = [(64, 1, 3, 3), (257,), (128, 64, 3, 3), (10, 256), (1,), (3, 5, 7), (40000,)]
    cpu = [torch.nn.Parameter(torch.randn(s, generator=g) * 0.1) for s in shapes]
    opt = torch.optim.Adam(cpu, lr=1e-3)
    ps = [p.detach().clone().to(dev) for p in cpu]
    ms, vs = [torch.zeros_like(p) for p in ps], [torch.zeros_like(p) for p in ps]
    gs = [torch.zeros_like(p) for p in ps]
    CH = 4096
    tp, tg, tm, tv, numel, ct, co, chunks = _adam_tables(ps, gs, ms, vs, dev, CH)
    step = torch.zeros(1, dtype=torch.int32, device=dev)
    scale = 0.25
    gsc = torch.full((1,), scale, device=dev)
    lr_dev = torch.zeros(1, device=dev)
    for it, lr in enumerate((1e-3, 3e-4, 7.5e-4)):
        for grp in opt.param_groups:
            grp["lr"] = lr
        for p, gd in zip(cpu, gs):
            graw = torch.randn(p.shape, generator=g) * (10.0 ** (it - 1))
            p.grad = graw * scale
            gd.copy_(graw)
        opt.step()
        L.check(lib.td_counter_add(step.data_ptr(), 1, L.stream_ptr()), "td_counter_add")
        if use_lr_dev:
            L.check(lib.td_fill_f32(lr_dev.data_ptr(), 1, lr, L.stream_ptr()), "td_fill_f32")
        L.check(lib.td_adam_multi(tp.data_ptr(), tg.data_ptr(), tm.data_ptr(), tv.data_ptr(), numel.data_ptr(), ct.data_ptr(),
                                  co.data_ptr(), chunks, CH, step.data_ptr(), (123.0 if use_lr_dev else lr),
                                  lr_dev.data_ptr() if use_lr_dev else None, 0.9, 0.999, 1e-8, gsc.data_ptr(), None,
                                  L.stream_ptr()), "td_adam_multi")
        for i, (p, q) in enumerate(zip(ps, cpu)):
            assert rel(p, q.detach()) < 1e-6, (it, i)
            st = opt.state[q]
            assert rel(ms[i], st["exp_avg"]) < 1e-6 and rel(vs[i], st["exp_avg_sq"]) < 1e-6, (it, i)


@pytest.mark.parametrize("n", [5, 1024, 11182273 // 4 * 4 + 360])
@pytest.mark.parametrize("max_norm,world", [(10.0, 1), (0.5, 1), (0.5, 4), (0.0, 2)])
def test_grad_clip_scale_vs_torch(dev, n, max_norm, world):
    from tinydiff import _lib as L
    lib = L.load()
    g = torch.Generator().manual_seed(n % 97)
    flat_sum = torch.randn(n, generator=g) * 0.01 * world           # what the all-reduce (sum) leaves in the buffer
    mean = [torch.nn.Parameter((flat_sum / world).clone())]
    mean[0].grad = mean[0].detach().clone()
    want_norm = float(torch.nn.utils.clip_grad_norm_(mean, max_norm if max_norm > 0 else float("inf")))
    d = flat_sum.to(dev)
    part = torch.zeros(int(lib.td_grad_clip_num_partials(n)), device=dev)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    scale, norm = torch.zeros(1, device=dev), torch.zeros(1, device=dev)
    for _ in range(2):                                               # the counter resets itself: a second launch works
        L.check(lib.td_grad_clip_scale(d.data_ptr(), n, 1.0 / world, max_norm, part.data_ptr(), cnt.data_ptr(),
                                       scale.data_ptr(), norm.data_ptr(), L.stream_ptr()), "td_grad_clip_scale")
    # against the norm in double: torch's own fp32 norm of 11 M elements is 4e-4 off (33.4256 vs 33.4391 measured), so
    # clip_grad_norm_'s result is only a loose cross-check at that size
    exact = float((flat_sum.double() / world).norm())
    assert abs(float(norm) - exact) / exact < 1e-6
    assert abs(float(norm) - want_norm) / want_norm < (1e-5 if n < 100000 else 1e-3)
    clip = min(1.0, max_norm / (exact + 1e-6)) if max_norm > 0 else 1.0
    assert rel(d * scale, flat_sum / world * clip) < 1e-6
    assert int(cnt) == 0


@pytest.mark.parametrize("name,sched", [("conditional_diffusion_laion", "per_batch"), ("diffusion_transformer", "per_epoch")])
def test_trainstep_clip_and_cosine_schedule_vs_torch(dev, name, sched):
    """The optimizer half of the reference train loops, unchanged lines on ``TrainStep.optimizer``:
    LAION -- ``clip_grad_norm_(10.0)`` + Adam(1e-4) + CosineAnnealingLR(T_max=num_epochs, eta_min=1e-6) stepped per batch
    (conditional_diffusion_laion.py:434-438,471-473); DiT -- Adam(3e-4) + CosineAnnealingLR(T_max=num_epochs) stepped per
    epoch (diffusion_transformer.py:176-177,288).  Each step's gradients (read back from the step) are fed to torch's own
    clip + Adam + scheduler on a CPU copy of the parameters: the parameters must agree to 1e-6 after every step, under the
    captured graph."""
    import importlib
    from tinydiff.train import TrainStep
    mod = importlib.import_module(f"tinydiff.{name}")
    torch.manual_seed(0)
    kw = {"dropout": 0.0} if name == "diffusion_transformer" else {}
    model = mod.NoiseModel(**kw)
    model.load_state_dict(init_state_dict(name), strict=True)
    model.precision = "fp32"
    model = model.to(dev).train()
    fp = mod.ForwardProcess()
    Bn = 6
    lr0 = 1e-4 if sched == "per_batch" else 3e-4
    # the reference clips at 10.0; the clip must also be exercised where it is active (the norm at init is well below 10)
    max_norm = 0.05 if sched == "per_batch" else None
    ts = TrainStep(model, fp, Bn, dev, lr=lr0, max_grad_norm=max_norm, use_graph=True)
    scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(ts.optimizer, T_max=3, **({"eta_min": 1e-6} if sched == "per_batch" else {}))
    cpu = {k: torch.nn.Parameter(p.detach().cpu().clone()) for k, p in model.named_parameters()}
    opt = torch.optim.Adam(list(cpu.values()), lr=lr0)
    ref_sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=3, **({"eta_min": 1e-6} if sched == "per_batch" else {}))
    clipped_once = False
    for step in range(5):
        inp = make_inputs(name, Bn, seed=300 + step)
        ts(inp["x0"], inp["cond"], t=inp["t"], noise=inp["noise"])
        torch.cuda.synchronize()
        for k, p in cpu.items():
            p.grad = ts.eng.pgrad[k].detach().cpu().clone()
        if max_norm is not None:
            total = float(torch.nn.utils.clip_grad_norm_(list(cpu.values()), max_norm))
            assert abs(float(ts.grad_norm) - total) / total < 1e-5
            clipped_once |= total > max_norm
        opt.step()
        for k, p in model.named_parameters():
            assert rel(p, cpu[k].detach()) < 1e-6, (step, k)
        if sched == "per_batch" or step % 2 == 1:                    # "epochs" of two batches for the per-epoch schedule
            scheduler.step()
            ref_sched.step()
        assert ts.optimizer.param_groups[0]["lr"] == opt.param_groups[0]["lr"]
    assert max_norm is None or clipped_once
    assert ts.lr != lr0


def test_trainstep_set_lr_under_graph(dev):
    """lr is a device scalar: changing it after the graph was captured changes the update (it was frozen at capture)."""
    from tinydiff.conditional_diffusion import ForwardProcess, NoiseModel
    from tinydiff.train import TrainStep
    name = "conditional_diffusion"
    inp = make_inputs(name, 4)
    fp = ForwardProcess()
    upd = []
    for lr2 in (1e-3, 0.0):
        m = NoiseModel()
        m.load_state_dict(init_state_dict(name), strict=True)
        m = m.to(dev).train()
        ts = TrainStep(m, fp, 4, dev, lr=1e-3, use_graph=True)
        ts(inp["x0"], inp["cond"], t=inp["t"], noise=inp["noise"])
        w1 = m.final_conv.weight.detach().clone()
        ts.set_lr(lr2)
        ts(inp["x0"], inp["cond"], t=inp["t"], noise=inp["noise"])
        upd.append(float((m.final_conv.weight.detach() - w1).abs().max()))
    assert upd[0] > 1e-5 and upd[1] == 0.0


def test_trainstep_device_rng_in_graph(dev):
    """Nothing injected: t ~ randint(0, T) and the q_sample noise are drawn inside the captured step (td_randint + Philox):
    seeded, different every step, in range, unit variance; no library RNG kernel involved."""
    from tinydiff.conditional_diffusion import ForwardProcess, NoiseModel
    from tinydiff.train import TrainStep
    name = "conditional_diffusion"
    Bn = 64
    inp = make_inputs(name, Bn)
    fp = ForwardProcess()

    def run(seed):
        m = NoiseModel()
        m.load_state_dict(init_state_dict(name), strict=True)
        m = m.to(dev).train()
        ts = TrainStep(m, fp, Bn, dev, seed=seed)
        out = []
        for _ in range(3):
            loss = float(ts(inp["x0"], inp["cond"]))
            out.append((loss, ts.eng.t_in.cpu().clone(), ts.noise.cpu().clone()))
        assert ts.launches_per_step > 0
        return out

    a, b, c = run(11), run(11), run(12)
    for (la, ta, na), (lb, tb, nb) in zip(a, b):
        assert la == lb and torch.equal(ta, tb) and torch.equal(na, nb)
    assert not torch.equal(a[0][1], c[0][1]) and not torch.equal(a[0][2], c[0][2])
    assert not torch.equal(a[0][1], a[1][1]) and not torch.equal(a[0][2], a[1][2])
    t_all = torch.cat([x[1] for x in a])
    assert int(t_all.min()) >= 0 and int(t_all.max()) < 1000 and t_all.unique().numel() > 100
    n_all = torch.cat([x[2].flatten() for x in a])
    assert abs(float(n_all.mean())) < 0.01 and abs(float(n_all.std()) - 1.0) < 0.01
    # x_t of the last step is q_sample of the drawn (t, noise)
    # (x_in still holds it: the forward does not overwrite its input)


def test_psample_out_of_range_step_is_noop_and_odd_noise_rows(dev):
    """td_psample_step: a step counter outside [0, T) leaves x untouched (a graph replayed past t = 0); a noise table
    whose rows are not 16-byte aligned (n % 4 != 0, e.g. an odd latent size) takes the scalar path and stays bit-exact."""
    from tinydiff import _lib as L
    from tinydiff.process import ForwardProcess
    lib = L.load()
    fp = ForwardProcess(num_timesteps=10)
    tab = fp._tables(dev)
    g = torch.Generator().manual_seed(2)
    n = 3 * 7                                            # 21 floats per noise row: rows 1, 2, 3 ... are misaligned
    x, eps = torch.randn(n, generator=g), torch.randn(n, generator=g)
    z = torch.randn(10, n, generator=g)
    zd, ed = z.to(dev), eps.to(dev)
    for t in (-1, 10, 11):
        xd = x.to(dev)
        td = torch.tensor([t], dtype=torch.int32, device=dev)
        L.check(lib.td_psample_step(xd.data_ptr(), ed.data_ptr(), zd.data_ptr(), n, tab["coef"].data_ptr(),
                                    td.data_ptr(), n, 10, None, L.stream_ptr()), "td_psample_step")
        assert torch.equal(xd.cpu(), x), t
    for t in (9, 5, 1, 0):
        xd = x.to(dev)
        td = torch.tensor([t], dtype=torch.int32, device=dev)
        L.check(lib.td_psample_step(xd.data_ptr(), ed.data_ptr(), zd.data_ptr(), n, tab["coef"].data_ptr(),
                                    td.data_ptr(), n, 10, None, L.stream_ptr()), "td_psample_step")
        want = O.p_sample_step(x, eps, z[t], t, fp.betas, fp.alphas, fp.alphas_cumprod)
        assert torch.equal(xd.cpu(), want), t
    from tinydiff.process import ReverseLoop
    loop = ReverseLoop(fp, x.to(dev), eps.to(dev), torch.zeros(1, dtype=torch.int32, device=dev), lambda: None, use_graph=False)
    with pytest.raises(ValueError):
        loop.run(z=None, seed=1, steps=11)


def test_psample_step_advance_counts_down(dev):
    """td_psample_step_advance == td_psample_step followed by t_dev -= 1 (the last block of the grid writes the counter), for
    tensors of one block and of many blocks, across the whole schedule and past its end (no-op steps still count down)."""
    from tinydiff import _lib as L
    from tinydiff.process import ForwardProcess
    lib = L.load()
    fp = ForwardProcess(num_timesteps=6)
    tab = fp._tables(dev)
    for n in (20, 2560, 128 * 784):
        g = torch.Generator().manual_seed(n)
        x0 = torch.randn(n, generator=g)
        eps = torch.randn(8, n, generator=g).to(dev)
        z = torch.randn(6, n, generator=g).to(dev)
        xa, xb = x0.to(dev).clone(), x0.to(dev).clone()
        ta = torch.tensor([5], device=dev, dtype=torch.int32)
        tb = ta.clone()
        ticket = torch.zeros(1, device=dev, dtype=torch.int32)
        for i in range(8):                     # two steps past t = 0
            L.check(lib.td_psample_step_advance(xa.data_ptr(), eps[i].data_ptr(), z.data_ptr(), n, tab["coef"].data_ptr(), ta.data_ptr(),
                                                n, 6, None, ticket.data_ptr(), L.stream_ptr()), "td_psample_step_advance")
            L.check(lib.td_psample_step(xb.data_ptr(), eps[i].data_ptr(), z.data_ptr(), n, tab["coef"].data_ptr(), tb.data_ptr(), n, 6,
                                        None, L.stream_ptr()), "td_psample_step")
            L.check(lib.td_counter_add(tb.data_ptr(), -1, L.stream_ptr()), "td_counter_add")
            assert int(ta.item()) == int(tb.item()) == 4 - i
            assert int(ticket.item()) == 0
            assert torch.equal(xa, xb)
