"""Kernel-level parity on a real B200: every C-ABI kernel against the CPU oracle / ATen fp32.
Bit-exact where the arithmetic is elementwise with the reference's rounding order; stated
tolerances otherwise."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import ddpm_oracle as O   # noqa: E402  (checker only)


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a B200"
    from tinydiff import _lib as L
    return L.require_device("cuda:0")


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("shape", [(5, 1, 28, 28), (7, 20), (3, 4, 32, 32), (1, 3)])
def test_qsample_bit_exact(dev, shape):
    from tinydiff import ops
    g = torch.Generator().manual_seed(1)
    _, _, ac = O.make_schedule()
    x0 = torch.randn(shape, generator=g)
    nz = torch.randn(shape, generator=g)
    t = torch.randint(0, 1000, (shape[0],), generator=g)
    want = O.q_sample(ac, x0, t, nz)
    got = ops.qsample(x0.to(dev), t.to(dev), nz.to(dev), ac.to(dev))
    assert torch.equal(got.cpu(), want)


def test_qsample_philox_statistics(dev):
    from tinydiff import ops
    _, _, ac = O.make_schedule()
    x0 = torch.zeros(64, 1, 128, 128, device=dev)
    t = torch.full((64,), 999, device=dev)
    x_t, nz = ops.qsample_philox(x0, t, ac.to(dev), seed=7)
    assert abs(float(nz.mean())) < 5e-3 and abs(float(nz.std()) - 1) < 5e-3
    assert abs(float((nz ** 4).mean()) - 3.0) < 0.05           # kurtosis of N(0,1)
    x_t2, nz2 = ops.qsample_philox(x0, t, ac.to(dev), seed=7)
    assert torch.equal(nz, nz2)
    _, nz3 = ops.qsample_philox(x0, t, ac.to(dev), seed=8)
    assert not torch.equal(nz, nz3)
    assert torch.equal(x_t.cpu(), O.q_sample(ac, x0.cpu(), t.cpu(), nz.cpu()))


@pytest.mark.parametrize("t", [0, 1, 500, 999])
@pytest.mark.parametrize("n", [2 * 784, 3 * 20 + 1])
def test_psample_bit_exact(dev, t, n):
    from tinydiff import ops
    from tinydiff.process import ForwardProcess
    fp = ForwardProcess()
    g = torch.Generator().manual_seed(2)
    x, eps, z = (torch.randn(n, generator=g) for _ in range(3))
    want = O.p_sample_step(x, eps, z, t, fp.betas, fp.alphas, fp.alphas_cumprod)
    xd = x.to(dev)
    ops.psample_step(xd, eps.to(dev), z.to(dev), fp._tables(dev)["coef"], t)
    assert torch.equal(xd.cpu(), want)


@pytest.mark.parametrize("n", [100352, 128 * 20, 17])
def test_mse_grad(dev, n):
    from tinydiff import ops
    g = torch.Generator().manual_seed(3)
    p, q = torch.randn(n, generator=g), torch.randn(n, generator=g)
    loss, grad = O.mse_loss_and_grad(p, q)
    l2, g2 = ops.mse_grad(p.to(dev), q.to(dev))
    assert abs(float(l2) - float(loss)) <= 2e-6 * float(loss)
    assert rel(g2, grad) < 1e-6
    l3, _ = ops.mse_grad(p.to(dev), q.to(dev))        # deterministic
    assert float(l3) == float(l2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("H,ceil", [(28, True), (7, True), (14, True), (32, False), (7, False)])
def test_maxpool(dev, dtype, H, ceil):
    from tinydiff import ops
    g = torch.Generator().manual_seed(4)
    x = torch.randn(3, 16, H, H, generator=g).to(dtype)
    want = F.max_pool2d(x.float(), 2, ceil_mode=ceil)
    got = nchw(ops.maxpool2(nhwc(x).to(dev), ceil)).float().cpu()
    assert torch.equal(got, want)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.bfloat16, 4e-3)])
@pytest.mark.parametrize("hl,hs", [(4, 7), (8, 14), (16, 28), (8, 16)])
def test_upcat(dev, dtype, tol, hl, hs):
    from tinydiff import ops
    g = torch.Generator().manual_seed(5)
    B, cu, cs = 3, 16, 24
    low = torch.randn(B, cu, hl, hl, generator=g).to(dtype)
    skip = torch.randn(B, cs, hs, hs, generator=g).to(dtype)
    temb = torch.randn(B, 40, generator=g)
    toff = 8
    up = F.interpolate(low.float(), scale_factor=2, mode="bilinear", align_corners=True)
    s = skip.float() + temb[:, toff:toff + cs].view(B, cs, 1, 1)
    if hs != 2 * hl:
        s = F.interpolate(s, size=(2 * hl, 2 * hl), mode="bilinear", align_corners=True)
    want = torch.cat([up, s], 1)
    got = nchw(ops.upcat(nhwc(low).to(dev), nhwc(skip).to(dev), temb.to(dev), toff)).float().cpu()
    assert rel(got, want) < tol


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.bfloat16, 4e-3)])
def test_resize(dev, dtype, tol):
    from tinydiff import ops
    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, 64, 32, 32, generator=g).to(dtype)
    want = F.interpolate(x.float(), size=(28, 28), mode="bilinear", align_corners=True)
    got = nchw(ops.resize_bilinear(nhwc(x).to(dev), 28, 28)).float().cpu()
    assert rel(got, want) < tol


@pytest.mark.parametrize("mode", ["time", "class", "sin_text"])
def test_embed_head(dev, mode):
    from tinydiff import ops
    g = torch.Generator().manual_seed(7)
    B = 6
    D = 768 if mode == "sin_text" else 256
    din = D if mode == "sin_text" else 1
    w0, b0 = torch.randn(D, din, generator=g) / din ** 0.5, torch.randn(D, generator=g) * 0.1
    w2, b2 = torch.randn(D, D, generator=g) / D ** 0.5, torch.randn(D, generator=g) * 0.1
    pw, pb = torch.randn(896, D, generator=g) / D ** 0.5, torch.randn(896, generator=g) * 0.1
    t = torch.randint(0, 1000, (B,), generator=g)
    y = torch.randint(0, 10, (B,), generator=g)
    table = torch.randn(10, D, generator=g)
    text = torch.randn(B, D, generator=g)
    if mode == "sin_text":
        e = O.timestep_embedding_sinusoidal(t, D)
    else:
        e = t.unsqueeze(-1).float()
    e = F.linear(F.silu(F.linear(e, w0, b0)), w2, b2)
    if mode == "class":
        e = e + table[y]
    if mode == "sin_text":
        e = e + text
    want = F.linear(e, pw, pb)
    d = lambda v: v.to(dev)
    got, emb = ops.embed_head(d(t), d(w0), d(b0), d(w2), d(b2), d(pw), d(pb), in_mode=2 if mode == "sin_text" else 0,
                              y=d(y) if mode == "class" else None, class_table=d(table) if mode == "class" else None,
                              text=d(text) if mode == "sin_text" else None)
    assert rel(emb, e) < 2e-5
    assert rel(got, want) < 2e-5


def _conv_case(dev, B, H, cin, cout, engine, in_dtype, out_dtype, relu=True, seed=0):
    from tinydiff import ops
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, cin, H, H, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) / (9 * cin) ** 0.5
    scale = 0.5 + torch.rand(cout, generator=g)
    shift = torch.randn(cout, generator=g) * 0.1
    wdt = torch.bfloat16 if engine == 1 else torch.float32
    xr = x.to(in_dtype).float()
    wr = w.to(wdt).float()
    want = F.conv2d(xr.double(), wr.double(), padding=1) * scale.double().view(1, -1, 1, 1) + shift.double().view(1, -1, 1, 1)
    if relu:
        want = want.clamp_min(0)
    wp = ops.pack_conv_weight(w.to(dev), wdt)
    got = ops.conv3x3(nhwc(x).to(in_dtype).to(dev), wp, scale.to(dev), shift.to(dev), relu, engine, out_dtype)
    return nchw(got).double().cpu(), want


@pytest.mark.parametrize("B,H,cin,cout", [(2, 28, 64, 128), (3, 7, 256, 512), (2, 5, 3, 10), (1, 32, 32, 64)])
def test_conv_simt_fp32(dev, B, H, cin, cout):
    got, want = _conv_case(dev, B, H, cin, cout, 0, torch.float32, torch.float32)
    assert rel(got, want) < 2e-6


# every (H, Cin, Cout) the MNIST UNet runs through the tensor-core engine at a
# small batch, plus ragged batches that leave partially filled M tiles
TC_SHAPES = [(28, 64, 128), (28, 128, 128), (14, 128, 256), (14, 256, 256), (7, 256, 512), (7, 512, 512),
             (4, 512, 512), (8, 1024, 256), (8, 256, 256), (16, 512, 128), (16, 128, 128), (32, 256, 64),
             (32, 64, 64)]


@pytest.mark.parametrize("H,cin,cout", TC_SHAPES)
@pytest.mark.parametrize("B", [3, 19])
def test_conv_tc_bf16(dev, B, H, cin, cout):
    got, want = _conv_case(dev, B, H, cin, cout, 1, torch.bfloat16, torch.bfloat16, seed=H + cin)
    # inputs are rounded to bf16 on both sides; the only differences are fp32 accumulation order
    # and the final bf16 rounding of the output (2^-9 relative per element)
    assert rel(got, want) < 3e-3
    got32, want = _conv_case(dev, B, H, cin, cout, 1, torch.bfloat16, torch.float32, seed=H + cin)
    assert rel(got32, want) < 2e-5


# The halo kernel (conv_halo.cu): both shared-memory layouts forced on every map size they accept, the
# per-tap kernel forced on the same shapes, and batches large enough that each persistent CTA walks
# several (n-tile, subtile) units with paired and unpaired subtiles.
@pytest.mark.parametrize("mode", ["1", "2", "3", "off"])
@pytest.mark.parametrize("B,H,cin,cout", [(5, 28, 64, 128), (3, 14, 128, 256), (7, 7, 256, 128), (2, 16, 128, 128),
                                          (3, 32, 64, 64), (2, 8, 64, 64), (9, 4, 64, 128), (66, 28, 64, 128),
                                          (150, 16, 64, 64), (33, 20, 64, 64), (5, 24, 64, 64), (40, 32, 128, 64)])
def test_conv_halo_layouts(dev, monkeypatch, mode, B, H, cin, cout):
    if mode == "off":
        monkeypatch.setenv("TD_TC_HALO", "0")
    else:
        monkeypatch.setenv("TD_TC_HALO_MODE", mode)
    got32, want = _conv_case(dev, B, H, cin, cout, 1, torch.bfloat16, torch.float32, seed=B + H)
    assert rel(got32, want) < 2e-5
    assert float((got32 - want).abs().max()) < 1e-3           # no stray element (halo / dead-column handling)


# MaxPool2d(2) fused into the halo epilogue (eval mode, bf16): every map the UNets pool (28 / 14 / 7 with ceil_mode, the latent
# UNet's 32 / 16 in strips) plus floor mode, ragged batches and the benchmark batch.  The pooled map must be BIT-identical to
# pooling the conv's own bf16 output (max commutes with the monotone bf16 rounding).
@pytest.mark.parametrize("B,H,cin,cout,mode", [(5, 28, 128, 128, "ceil"), (3, 14, 256, 256, "ceil"), (8, 7, 512, 512, "ceil"),
                                               (128, 28, 64, 128, "ceil"), (37, 14, 128, 256, "ceil"), (129, 7, 256, 512, "ceil"),
                                               (3, 32, 64, 64, "floor"), (5, 16, 64, 128, "ceil"), (4, 7, 64, 64, "floor"),
                                               (2, 20, 64, 64, "ceil"), (5, 24, 64, 128, "ceil")])
def test_conv_halo_fused_maxpool(dev, B, H, cin, cout, mode):
    from tinydiff import ops
    g = torch.Generator().manual_seed(B * 100 + H)
    x = torch.randn(B, H, H, cin, generator=g).to(dev).to(torch.bfloat16)
    w = ops.pack_conv_weight((torch.randn(cout, cin, 3, 3, generator=g) / (9 * cin) ** 0.5).to(dev), torch.bfloat16)
    scale = (torch.rand(cout, generator=g) + 0.5).to(dev)
    shift = (torch.randn(cout, generator=g) * 0.3).to(dev)
    for relu in (True, False):
        y_plain = ops.conv3x3(x, w, scale, shift, relu, 1, torch.bfloat16)
        y, pooled = ops.conv3x3(x, w, scale, shift, relu, 1, torch.bfloat16, pool=mode)
        assert torch.equal(y, y_plain), "the un-pooled output must not change"
        assert pooled is not None, "the halo plan must accept the fused pool on these maps"
        want = F.max_pool2d(y.float().permute(0, 3, 1, 2), 2, ceil_mode=(mode == "ceil")).permute(0, 2, 3, 1)
        assert pooled.shape == want.shape
        assert torch.equal(pooled.float(), want), f"relu={relu}: max abs diff {float((pooled.float() - want).abs().max())}"


def test_conv_fused_maxpool_declined(dev):
    """Plans that cannot pool in the epilogue (per-tap kernel on the 4x4 map, fp32 output) say so and leave pool_y untouched."""
    from tinydiff import ops
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 4, 4, 64, generator=g).to(dev).to(torch.bfloat16)
    w = ops.pack_conv_weight((torch.randn(64, 64, 3, 3, generator=g) / 24.0).to(dev), torch.bfloat16)
    _, pooled = ops.conv3x3(x, w, None, None, True, 1, torch.float32, pool="ceil")
    assert pooled is None


def test_conv_halo_channel_slices(dev):
    """Input read from / output written into a channel slice of a wider NHWC buffer (the decoder concat)."""
    from tinydiff import ops
    g = torch.Generator().manual_seed(5)
    B, H, cin, cout = 4, 28, 64, 128
    xw = torch.randn(B, H, H, 192, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) / (9 * cin) ** 0.5
    out = torch.full((B, H, H, 256), 7.0, device=dev, dtype=torch.float32)
    ops.conv3x3(xw.to(dev).to(torch.bfloat16), ops.pack_conv_weight(w.to(dev), torch.bfloat16), None, None, False, 1,
                torch.float32, out=out, y_coff=128, x_coff=64, cin=cin)
    xr = xw[..., 64:128].to(torch.bfloat16).double().permute(0, 3, 1, 2)
    want = F.conv2d(xr, w.to(torch.bfloat16).double(), padding=1)
    assert rel(nchw(out[..., 128:]), want) < 2e-5
    assert float((out[..., :128] - 7.0).abs().max()) == 0.0


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 3e-6), (torch.bfloat16, 3e-6)])
@pytest.mark.parametrize("B,hi,ho,c", [(5, 32, 28, 64), (130, 32, 28, 64), (3, 28, 28, 64), (2, 16, 20, 32), (1, 9, 7, 8)])
def test_final_resize_conv(dev, dtype, tol, B, hi, ho, c):
    """interpolate(align_corners=True) + final_conv (C -> 1) fused; inputs rounded to the activation dtype on both sides."""
    from tinydiff import ops
    g = torch.Generator().manual_seed(B + hi + c)
    x = torch.randn(B, c, hi, hi, generator=g).to(dtype).float()
    w = torch.randn(1, c, 3, 3, generator=g) / (9 * c) ** 0.5
    b = torch.randn(1, generator=g)
    want = F.conv2d(F.interpolate(x.double(), size=(ho, ho), mode="bilinear", align_corners=True), w.double(), b.double(), padding=1)
    got = ops.final_resize_conv(nhwc(x).to(dtype).to(dev), ops.pack_conv_weight(w.to(dev)), b.to(dev), ho, ho)
    assert rel(got, want) < tol


@pytest.mark.parametrize("B,H,W", [(3, 32, 32), (2, 7, 9), (5, 33, 17), (1, 1, 1), (2, 40, 40), (2, 3, 64)])
def test_conv_direct_latent_boundary(dev, B, H, W):
    """The 4-channel boundary layers of the latent UNet: initial_conv 4 -> 32 (two pixels per thread, written into the
    lower half of a 64-channel buffer) and final_conv 64 / 128 -> 4 (contract-then-stencil, several row bands)."""
    from tinydiff import ops
    g = torch.Generator().manual_seed(B * 100 + H + W)
    x = torch.randn(B, 4, H, W, generator=g)
    w = torch.randn(32, 4, 3, 3, generator=g) / 6
    b = torch.randn(32, generator=g) * 0.1
    want = F.conv2d(x.double(), w.double(), b.double(), padding=1)
    for odt, tol in ((torch.float32, 2e-6), (torch.bfloat16, 4e-3)):
        out = torch.full((B, H, W, 64), 3.0, device=dev, dtype=odt)
        ops.conv3x3(x.to(dev), ops.pack_conv_weight(w.to(dev)), None, b.to(dev), False, 2, odt, x_nchw=True, out=out)
        assert rel(nchw(out[..., :32]).double(), want) < tol
        assert float((out[..., 32:].float() - 3.0).abs().max()) == 0.0
    w64 = torch.randn(64, 4, 3, 3, generator=g) / 6            # final_conv's data gradient: 4 -> 64, two channel slices
    got = ops.conv3x3(x.to(dev), ops.pack_conv_weight(w64.to(dev)), None, None, False, 2, torch.float32, x_nchw=True)
    assert rel(nchw(got).double(), F.conv2d(x.double(), w64.double(), padding=1)) < 2e-6
    for cin in (64, 128):
        xa = torch.randn(B, cin, H, W, generator=g)
        w2 = torch.randn(4, cin, 3, 3, generator=g) / (9 * cin) ** 0.5
        b2 = torch.randn(4, generator=g)
        for idt in (torch.float32, torch.bfloat16):
            xr = xa.to(idt)
            want2 = F.conv2d(xr.double(), w2.double(), b2.double(), padding=1)
            wide = torch.zeros(B, H, W, cin + 8, dtype=idt)
            wide[..., 8:] = nhwc(xr)
            for y_nchw in (True, False):
                got2 = ops.conv3x3(wide.to(dev), ops.pack_conv_weight(w2.to(dev)), None, b2.to(dev), False, 2,
                                   torch.float32, y_nchw=y_nchw, x_coff=8, cin=cin)
                got2 = got2 if y_nchw else nchw(got2)
                # bf16 activations: mma.sync with the fp32 weights split into bf16 hi + lo (2^-17 relative per weight)
                assert rel(got2.double(), want2) < (2e-6 if idt == torch.float32 else 8e-6)


def test_conv_direct_first_last(dev):
    from tinydiff import ops
    g = torch.Generator().manual_seed(9)
    x = torch.randn(4, 1, 28, 28, generator=g)
    w = torch.randn(64, 1, 3, 3, generator=g) / 3
    b = torch.randn(64, generator=g) * 0.1
    want = F.conv2d(x, w, b, padding=1)
    for odt, tol in ((torch.float32, 2e-6), (torch.bfloat16, 4e-3)):
        got = ops.conv3x3(x.to(dev), ops.pack_conv_weight(w.to(dev)), None, b.to(dev), False, 2, odt, x_nchw=True)
        assert rel(nchw(got).float(), want) < tol
    for cout, cin in ((1, 64), (4, 64)):
        xa = torch.randn(3, cin, 28, 28, generator=g)
        w2 = torch.randn(cout, cin, 3, 3, generator=g) / (9 * cin) ** 0.5
        b2 = torch.randn(cout, generator=g)
        for idt, tol in ((torch.float32, 2e-6), (torch.bfloat16, 1e-6)):
            xr = xa.to(idt)
            want2 = F.conv2d(xr.float(), w2, b2, padding=1)
            got2 = ops.conv3x3(nhwc(xr).to(dev), ops.pack_conv_weight(w2.to(dev)), None, b2.to(dev), False, 2,
                               torch.float32, y_nchw=True)
            assert rel(got2, want2) < max(tol, 2e-6)
