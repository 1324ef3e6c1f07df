"""CPU-side checks of the C-ABI boundary: the shared library loads without a GPU, exports every
symbol include/tinydiff.h declares, the ctypes table in tinydiff/_lib.py has one row per declaration
with the same argument count, and the product path refuses to run without a B200 (no CPU fallback)."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declarations():
    h = open(os.path.join(ROOT, "include", "tinydiff.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    out = {}
    for name, args in re.findall(r"\n(?:int64_t|int|void|double|const char\*)\s+(td_\w+)\s*\(([^;]*?)\)\s*;", h):
        out[name] = 0 if args.strip() in ("void", "") else len(args.split(","))
    return out


def test_library_exports_every_declared_symbol():
    from tinydiff import _lib
    lib = _lib.load()
    decl = _declarations()
    assert len(decl) >= 45
    for name, nargs in decl.items():
        assert hasattr(lib, name), f"{name} declared in tinydiff.h but not exported"
        assert name in _lib._SIGS, f"{name} has no ctypes signature"
        assert len(_lib._SIGS[name][1]) == nargs, f"{name}: header has {nargs} args, ctypes {len(_lib._SIGS[name][1])}"
    for name in _lib._SIGS:
        assert name in decl, f"{name} bound in _lib.py but not declared in tinydiff.h"
    assert lib.td_version() >= 100


def test_struct_sizes_match_header_layout():
    """ctypes mirrors of the argument structs: field order / count against the header text."""
    from tinydiff import _lib
    h = open(os.path.join(ROOT, "include", "tinydiff.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    for cname, cls in (("td_embed_args", _lib.EmbedArgs), ("td_embed_grads", _lib.EmbedGrads),
                       ("td_conv3x3_desc", _lib.ConvDesc), ("td_wgrad_desc", _lib.WgradDesc),
                       ("td_gemm_args", _lib.GemmArgs)):
        body = re.search(r"typedef struct \{([^{}]*)\}\s*" + cname + ";", h, flags=re.S).group(1)
        fields = []
        for stmt in body.split(";"):
            stmt = stmt.strip()
            if not stmt:
                continue
            names = [re.sub(r"[\s\*]", "", part.split()[-1]) if " " in part.strip() else part.strip().lstrip("*")
                     for part in stmt.split(",")]
            fields += [n.lstrip("*") for n in names]
        got = [f[0] for f in cls._fields_]
        # proj_out_ptr etc. keep their names; compare the sequences
        assert got == fields, (cname, got, fields)


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box without a GPU")
def test_no_cpu_fallback():
    from tinydiff import _lib
    from tinydiff.diffusion import ForwardProcess, NoiseModel, sample
    with pytest.raises(RuntimeError):
        _lib.require_device("cpu")
    fp = ForwardProcess()
    with pytest.raises(RuntimeError):
        fp.q_sample("cpu", torch.zeros(2, 1, 28, 28), torch.zeros(2, dtype=torch.long))
    m = NoiseModel()
    with pytest.raises(RuntimeError):
        m.eval()(torch.zeros(2, 1, 28, 28), torch.zeros(2, dtype=torch.long))
    with pytest.raises(RuntimeError):
        sample(m, fp, "cpu", n_samples=2)
    # the C entry points themselves refuse as well
    lib = _lib.load()
    assert lib.td_device_check(0) != 0
    assert b"" != lib.td_last_error_string()


def test_process_attributes_match_reference_contract():
    from tinydiff.process import ForwardProcess
    fp = ForwardProcess()
    assert fp.num_timesteps == 1000
    for t in (fp.betas, fp.alphas, fp.alphas_cumprod):
        assert t.dtype == torch.float32 and t.device.type == "cpu" and t.shape == (1000,)
    assert float(fp.betas[0]) == pytest.approx(9.999999747e-05, rel=1e-7)
    assert float(fp.alphas_cumprod[-1]) == pytest.approx(4.035830e-05, rel=1e-5)
    assert fp.alphas[500].dim() == 0


def test_state_dict_layout_matches_fixture_modules():
    """state_dict keys / shapes / dtypes identical to the reference layout."""
    from oracle.fixtures import init_state_dict
    import importlib
    for name in ("diffusion", "conditional_diffusion", "conditional_diffusion_laion", "latent_diffusion",
                 "diffusion_transformer"):
        mod = importlib.import_module(f"tinydiff.{name}")
        model = mod.NoiseModel()
        ref = init_state_dict(name, perturb=(name != "diffusion_transformer"))
        mine = model.state_dict()
        assert list(mine.keys()) == list(ref.keys()), name
        for k in ref:
            assert mine[k].shape == ref[k].shape and mine[k].dtype == ref[k].dtype, (name, k)
        model.load_state_dict(ref, strict=True)
