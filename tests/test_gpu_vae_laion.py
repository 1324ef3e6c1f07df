"""SURVEY.md 8f #3 on the device: the vae_laion kernels (stride-2 4x4 convolution, ConvTranspose2d, flash-style
SelfAttention, spectral-norm sigma) against ATen / the oracle, and the drop-in VAE.encode / decode against the oracle and
the reference-generated golden outputs.  fp32 kernels: 1e-4 (attention uses ex2-based exponentials: 1e-4 on the output)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import vae_laion_oracle as V                 # noqa: E402  (checker only)
from oracle.fixtures import checksum, init_state_dict    # noqa: E402


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    from tinydiff import _lib as L
    return L.require_device("cuda:0")


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("B,H,cin,cout,x_nchw,act", [(2, 16, 3, 32, True, 1), (3, 8, 32, 64, False, 1), (1, 12, 20, 70, False, 0),
                                                      (2, 32, 128, 256, False, 1)])
def test_conv4x4s2(dev, B, H, cin, cout, x_nchw, act):
    from tinydiff import _lib as L
    lib = L.load()
    g = torch.Generator().manual_seed(H + cin)
    x = torch.randn(B, cin, H, H, generator=g)
    w = torch.randn(cout, cin, 4, 4, generator=g) / (16 * cin) ** 0.5
    b = torch.randn(cout, generator=g)
    sig = torch.tensor([1.7])
    want = F.conv2d(x, w / sig, b, stride=2, padding=1)
    want = F.relu(want) if act == 1 else want
    wp = torch.empty(cout * 16 * cin, device=dev)
    wd, sd_, bd = w.to(dev), sig.to(dev), b.to(dev)              # named: a temporary's block may be re-used before the launch
    L.check(lib.td_pack_conv4x4_weight(wd.data_ptr(), sd_.data_ptr(), wp.data_ptr(), cout, cin, 0, L.stream_ptr()))
    xd = (x if x_nchw else nhwc(x)).to(dev).contiguous()
    y = torch.empty(B, H // 2, H // 2, cout, device=dev)
    L.check(lib.td_conv4x4s2_fwd(xd.data_ptr(), wp.data_ptr(), bd.data_ptr(), y.data_ptr(), B, H, H, cin, cout, int(x_nchw),
                                 act, L.stream_ptr()))
    assert rel(nchw(y), want) < 1e-5


@pytest.mark.parametrize("B,H,cin,cout,y_nchw,act", [(2, 4, 256, 128, False, 1), (3, 8, 32, 3, True, 4), (1, 6, 20, 70, False, 0),
                                                      (2, 16, 64, 32, False, 1)])
def test_conv_transpose4x4s2(dev, B, H, cin, cout, y_nchw, act):
    from tinydiff import _lib as L
    lib = L.load()
    g = torch.Generator().manual_seed(H + cin)
    x = torch.randn(B, cin, H, H, generator=g)
    w = torch.randn(cin, cout, 4, 4, generator=g) / (4 * cin) ** 0.5
    b = torch.randn(cout, generator=g)
    sig = torch.tensor([0.6])
    want = F.conv_transpose2d(x, w / sig, b, stride=2, padding=1)
    want = F.relu(want) if act == 1 else (torch.sigmoid(want) if act == 4 else want)
    wp = torch.empty(4 * cout * 4 * cin, device=dev)
    wd, sd_, bd, xd = w.to(dev), sig.to(dev), b.to(dev), nhwc(x).to(dev)
    L.check(lib.td_pack_conv4x4_weight(wd.data_ptr(), sd_.data_ptr(), wp.data_ptr(), cout, cin, 1, L.stream_ptr()))
    y = torch.empty((B, cout, 2 * H, 2 * H) if y_nchw else (B, 2 * H, 2 * H, cout), device=dev)
    L.check(lib.td_convT4x4s2_fwd(xd.data_ptr(), wp.data_ptr(), bd.data_ptr(), y.data_ptr(), B, H, H, cin, cout,
                                  int(y_nchw), act, L.stream_ptr()))
    got = y if y_nchw else nchw(y)
    assert rel(got, want) < 1e-5


@pytest.mark.parametrize("C,H", [(32, 32), (64, 16), (128, 32), (32, 64)])
def test_self_attention_flash(dev, C, H):
    """vae_laion.SelfAttention.forward against the full-matrix computation (oracle), gamma != 0, logits with a wide range."""
    from tinydiff import _lib as L
    lib = L.load()
    g = torch.Generator().manual_seed(C + H)
    B = 2
    x = torch.randn(B, C, H, H, generator=g)
    p = "a"
    sd = {f"{p}.query.weight": torch.randn(C // 8, C, 1, 1, generator=g) * 0.3, f"{p}.query.bias": torch.randn(C // 8, generator=g) * 0.1,
          f"{p}.key.weight": torch.randn(C // 8, C, 1, 1, generator=g) * 0.3, f"{p}.key.bias": torch.randn(C // 8, generator=g) * 0.1,
          f"{p}.value.weight": torch.randn(C, C, 1, 1, generator=g) * 0.2, f"{p}.value.bias": torch.randn(C, generator=g) * 0.1,
          f"{p}.gamma": torch.tensor([0.8])}
    want = V.self_attention(sd, p, x)
    n, dq = H * H, C // 8
    xf = nhwc(x).view(B * n, C)
    wqkv = torch.cat([sd[f"{p}.query.weight"].view(-1, C), sd[f"{p}.key.weight"].view(-1, C), sd[f"{p}.value.weight"].view(-1, C)])
    bqkv = torch.cat([sd[f"{p}.query.bias"], sd[f"{p}.key.bias"], sd[f"{p}.value.bias"]])
    qkv = (xf @ wqkv.t() + bqkv).contiguous().to(dev)
    xd = xf.contiguous().to(dev)
    y = torch.empty_like(xd)
    gd = sd[f"{p}.gamma"].to(dev)
    L.check(lib.td_self_attention_fwd(qkv.data_ptr(), xd.data_ptr(), gd.data_ptr(), y.data_ptr(), B, n, dq, C, L.stream_ptr()))
    assert rel(nchw(y.view(B, H, H, C)), want) < 1e-4


@pytest.mark.parametrize("layer,dim1", [("encoder.1.0", False), ("decoder.0.0", True), ("encoder.3.2.conv1", False)])
@pytest.mark.parametrize("iters", [0, 1])
def test_spectral_sigma(dev, layer, dim1, iters):
    from tinydiff import _lib as L
    lib = L.load()
    sd = init_state_dict("vae_laion", perturb=False)           # raw (random) u, v: the power iteration has something to do
    w = sd[layer + ".weight_orig"]
    u, v = sd[layer + ".weight_u"].clone(), sd[layer + ".weight_v"].clone()
    new = {}
    wn = V.spectral_weight(sd, layer, dim=1 if dim1 else 0, training=bool(iters), new_uv=new)
    sigma_ref = float((w / wn).flatten()[w.flatten().abs().argmax()])
    rows = w.shape[1] if dim1 else w.shape[0]
    ud, vd, wd = u.to(dev), v.to(dev), w.to(dev)
    sig = torch.empty(1, device=dev)
    scratch = torch.empty(rows, device=dev)
    L.check(lib.td_spectral_sigma(wd.data_ptr(), rows, w.numel() // rows, int(dim1), w.shape[2] * w.shape[3], ud.data_ptr(),
                                  vd.data_ptr(), iters, 1e-12, sig.data_ptr(), scratch.data_ptr(), L.stream_ptr()))
    assert abs(float(sig) - sigma_ref) / abs(sigma_ref) < 1e-4
    if iters:
        assert rel(ud, new[layer + ".weight_u"]) < 1e-5 and rel(vd, new[layer + ".weight_v"]) < 1e-5
    else:
        assert torch.equal(ud.cpu(), u) and torch.equal(vd.cpu(), v)


def test_vae_encode_decode_vs_oracle_and_golden(dev, golden):
    from tinydiff.vae_laion import VAE
    g = golden("vae_laion")
    sd = init_state_dict("vae_laion")
    m = VAE()
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).eval()
    x = torch.rand(1, 3, 256, 256, generator=torch.Generator().manual_seed(77))
    z = torch.randn(1, 128, generator=torch.Generator().manual_seed(78))
    mu, logvar = m.encode(x.to(dev))
    rec = m.decode(z.to(dev))
    assert rel(mu, g["mu"]) < 1e-4 and rel(logvar, g["logvar"]) < 1e-4
    assert rel(rec[:, :, ::8, ::8], g["recon_sub"]) < 1e-4
    assert rel(checksum(rec.cpu()), g["recon_checksum"]) < 1e-4
    # a batch of 2 with fresh inputs against the oracle, and forward() with an injected reparameterisation
    x2 = torch.rand(2, 3, 256, 256, generator=torch.Generator().manual_seed(5))
    mu_o, lv_o = V.encode(sd, x2)
    mu2, lv2 = m.encode(x2.to(dev))
    assert rel(mu2, mu_o) < 1e-4 and rel(lv2, lv_o) < 1e-4
    eps = torch.randn(2, 128, generator=torch.Generator().manual_seed(6))
    z2 = m.reparameterize(mu2, lv2, eps=eps.to(dev))
    assert rel(z2, V.reparameterize(mu_o, lv_o, eps)) < 1e-4
    assert rel(m.decode(z2), V.decode(sd, z2.cpu())) < 1e-4
    with pytest.raises(NotImplementedError):
        m.train().encode(x.to(dev))
