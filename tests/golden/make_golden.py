"""Generate tests/golden/*.pt by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only:   python tests/golden/make_golden.py
The fixtures are committed; this script is committed beside them so they can be regenerated.
Weights are the reference's own default init under ``torch.manual_seed(0)`` (drawn after the
module import), with BatchNorm affine/running statistics perturbed by ``perturb_bn`` so that
eval-mode BatchNorm is not the identity.  Inputs come from ``torch.Generator().manual_seed(1234)``.
Only small tensors are stored (outputs, checksums); weights are re-created from the seed.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_shim as shim          # noqa: E402
from oracle.fixtures import (init_state_dict, build_reference_model, make_inputs, perturb_bn, checksum,   # noqa: E402
                             injected_randn)

OUT = os.path.dirname(os.path.abspath(__file__))


def unet_case(modname: str, B: int = 4, n_samp: int = 2):
    ref = shim.load(modname)
    model = build_reference_model(ref, modname)
    perturb_bn(model)
    fp = ref.ForwardProcess()
    inp = make_inputs(modname, B)
    cond = inp.get("cond")
    args = (lambda x, t: (x, t) if cond is None else (x, t, cond))
    g = {"schedule": {"betas": fp.betas.clone(), "alphas": fp.alphas.clone(),
                      "alphas_cumprod": fp.alphas_cumprod.clone()}}

    # q_sample with the reference's own arithmetic, noise injected through torch.randn_like
    with injected_randn(ref, [inp["noise"]]):
        x_t, noise = fp.q_sample(torch.device("cpu"), inp["x0"], inp["t"])
    assert torch.equal(noise, inp["noise"])
    g["x_t"] = x_t.clone()

    # eval-mode eps
    model.eval()
    with torch.no_grad():
        g["eps_eval"] = model(*args(x_t, inp["t"])).clone()

    # train step: loss, grads, BN buffers, one Adam step (diffusion.py:220-236)
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    pred = model(*args(x_t, inp["t"]))
    loss = torch.nn.functional.mse_loss(pred, noise)
    opt.zero_grad()
    loss.backward()
    g["eps_train"] = pred.detach().clone()
    g["loss"] = loss.detach().clone()
    g["grad_checksums"] = {k: checksum(p.grad) for k, p in model.named_parameters()}
    # a few full gradients for tight checks (small ones)
    g["grads_small"] = {k: p.grad.clone() for k, p in model.named_parameters()
                        if p.numel() <= 1024}
    opt.step()
    g["param_checksums_after_step"] = {k: checksum(v) for k, v in model.state_dict().items()}

    # reverse loop, n_samp samples, full T, injected x_T and z (diffusion.py:254-276)
    gen = torch.Generator().manual_seed(77)
    shape = (n_samp,) + tuple(inp["x0"].shape[1:])
    x_T = torch.randn(shape, generator=gen)
    T = fp.num_timesteps
    zs = [torch.randn(shape, generator=gen) for _ in range(T - 1)]     # used at t = T-1 .. 1
    kept = {}
    scond = None if cond is None else cond[:n_samp]

    model2 = build_reference_model(ref, modname)
    perturb_bn(model2)
    orig_forward = model2.forward

    def spy(x, t, *rest):
        out = orig_forward(x, t, *rest)
        tt = int(t[0])
        if tt % 100 == 0 or tt == T - 1:
            kept[tt] = (x.clone(), out.clone())
        return out
    model2.forward = spy
    with injected_randn(ref, [x_T] + zs):
        if modname == "diffusion":
            x0 = ref.sample(model2, fp, torch.device("cpu"), n_samples=n_samp)
        elif modname == "conditional_diffusion":
            x0 = ref.sample(model2, fp, torch.device("cpu"), n_samples=n_samp, y=scond)
    g["sample"] = {"x_0": x0.clone(), "kept": kept, "seed": 77, "n": n_samp}
    torch.save(g, os.path.join(OUT, f"{modname}.pt"))
    print(modname, "loss", float(loss), "eps_eval norm", float(g["eps_eval"].norm()),
          "x0 norm", float(x0.norm()))


def laion_case(B: int = 2):
    modname = "conditional_diffusion_laion"
    ref = shim.load(modname)
    model = build_reference_model(ref, modname)
    perturb_bn(model)
    fp = ref.ForwardProcess()
    inp = make_inputs(modname, B)
    g = {}
    with injected_randn(ref, [inp["noise"]]):
        x_t, noise = fp.q_sample(torch.device("cpu"), inp["x0"], inp["t"])
    g["x_t"] = x_t.clone()
    model.eval()
    with torch.no_grad():
        g["eps_eval"] = model(x_t, inp["t"], inp["cond"]).clone()
    model.train()
    pred = model(x_t, inp["t"], inp["cond"])
    loss = torch.nn.functional.mse_loss(pred, noise)
    loss.backward()
    g["eps_train"] = pred.detach().clone()
    g["loss"] = loss.detach().clone()
    g["grad_checksums"] = {k: checksum(p.grad) for k, p in model.named_parameters()}
    g["sin_emb"] = ref.get_timestep_embedding(inp["t"], 768).clone()
    torch.save(g, os.path.join(OUT, f"{modname}.pt"))
    print(modname, "loss", float(loss))


def dit_case(B: int = 8):
    modname = "diffusion_transformer"
    ref = shim.load(modname)
    model = build_reference_model(ref, modname)
    fp = ref.ForwardProcess()
    inp = make_inputs(modname, B)
    g = {}
    with injected_randn(ref, [inp["noise"]]):
        x_t, noise = fp.q_sample(torch.device("cpu"), inp["x0"], inp["t"])
    g["x_t"] = x_t.clone()
    model.eval()
    with torch.no_grad():
        g["eps_eval"] = model(x_t, inp["t"], inp["cond"]).clone()
    model.train()      # dropout=0.0 model: train == eval numerically
    pred = model(x_t, inp["t"], inp["cond"])
    loss = torch.nn.functional.mse_loss(pred, noise)
    loss.backward()
    g["loss"] = loss.detach().clone()
    g["grad_checksums"] = {k: checksum(p.grad) for k, p in model.named_parameters()}
    torch.save(g, os.path.join(OUT, f"{modname}.pt"))
    print(modname, "loss", float(loss))


def mlp_case(B: int = 8):
    modname = "latent_diffusion"
    ref = shim.load(modname)
    model = build_reference_model(ref, modname)
    perturb_bn(model)
    fp = ref.ForwardProcess()
    inp = make_inputs(modname, B)
    g = {}
    with injected_randn(ref, [inp["noise"]]):
        x_t, noise = fp.q_sample(torch.device("cpu"), inp["x0"], inp["t"])
    g["x_t"] = x_t.clone()
    model.eval()
    with torch.no_grad():
        g["eps_eval"] = model(x_t, inp["t"], inp["cond"]).clone()
    model.train()
    pred = model(x_t, inp["t"], inp["cond"])
    loss = torch.nn.functional.mse_loss(pred, noise)
    loss.backward()
    g["eps_train"] = pred.detach().clone()
    g["loss"] = loss.detach().clone()
    g["grad_checksums"] = {k: checksum(p.grad) for k, p in model.named_parameters()}
    g["buffers_after"] = {k: v.clone() for k, v in model.state_dict().items() if "running" in k or "tracked" in k}
    # VAE edges (vae.py:51-62) with seeded weights
    vae_mod = shim.load("vae")
    state = torch.get_rng_state()
    torch.manual_seed(0)
    vae = vae_mod.VAE(vae_mod.VAEConfig())
    torch.set_rng_state(state)
    xs = torch.rand(4, 784, generator=torch.Generator().manual_seed(9))
    with torch.no_grad():
        mu, logvar = vae.encode(xs)
        g["vae"] = {"x": xs, "mu": mu.clone(), "logvar": logvar.clone(), "dec": vae.decode(mu).clone()}
    fix = init_state_dict("vae", perturb=False)          # the fixture init reproduces VAE() under seed 0
    assert all(torch.equal(fix[k], v) for k, v in vae.state_dict().items())
    torch.save(g, os.path.join(OUT, f"{modname}.pt"))
    print(modname, "loss", float(loss))


def vae_laion_case():
    """vae_laion.VAE encode / decode (vae_laion.py:177-196) of the reference's own class (imported through
    shim.load_vae_laion: HF dataset, VGG16 download and the hard-coded CUDA device stubbed), fixture weights."""
    from oracle.fixtures import checksum
    mod = shim.load_vae_laion()
    sd = init_state_dict("vae_laion")
    model = shim.build_vae_laion(mod, sd)
    x = torch.rand(1, 3, 256, 256, generator=torch.Generator().manual_seed(77))
    z = torch.randn(1, 128, generator=torch.Generator().manual_seed(78))
    with torch.no_grad():
        mu, logvar = model.encode(x)
        rec = model.decode(z)
    g = {"x_seed": 77, "z_seed": 78, "mu": mu.clone(), "logvar": logvar.clone(), "recon_checksum": checksum(rec),
         "recon_sub": rec[:, :, ::8, ::8].clone()}
    # spectral_norm in a TRAINING forward: one power iteration of (u, v), then sigma (vae_laion.py:98 through torch's hook)
    conv = model.encoder[1][0]
    conv.train()
    u0, v0 = conv.weight_u.clone(), conv.weight_v.clone()
    conv(torch.zeros(1, 32, 8, 8))
    w = conv.weight_orig.detach()
    sigma = torch.dot(conv.weight_u, torch.mv(w.reshape(w.shape[0], -1), conv.weight_v))
    g["power_iter"] = {"layer": "encoder.1.0", "u0": u0, "v0": v0, "u1": conv.weight_u.clone(), "v1": conv.weight_v.clone(),
                       "sigma": sigma.clone()}
    torch.save(g, os.path.join(OUT, "vae_laion.pt"))
    print("vae_laion", float(mu.abs().mean()), float(rec.mean()))


if __name__ == "__main__":
    assert shim.available(), "reference not mounted"
    torch.set_num_threads(os.cpu_count() or 1)
    unet_case("diffusion")
    unet_case("conditional_diffusion")
    laion_case()
    dit_case()
    mlp_case()
    vae_laion_case()
