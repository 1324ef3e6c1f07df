"""Checkpoint compatibility: state_dict files written by the reference scripts load into the
drop-in modules, including the ``_orig_mod.`` key prefix of the torch.compile-wrapped LAION model
(conditional_diffusion_laion.py:527,614).  CPU only: no kernels run."""
import torch

from oracle.fixtures import init_state_dict


def _roundtrip(tmp_path, model, sd, prefix):
    path = tmp_path / "best_model.pth"
    torch.save({prefix + k: v for k, v in sd.items()}, path)        # what the reference's train() writes
    from tinydiff.checkpoint import load_checkpoint
    res = load_checkpoint(model, str(path))
    assert not res.missing_keys and not res.unexpected_keys
    got = model.state_dict()
    assert list(got) == list(sd), "state_dict key order differs from the reference layout"
    for k in sd:
        assert got[k].dtype == sd[k].dtype and torch.equal(got[k], sd[k]), k
    return got


def test_plain_checkpoint_loads(tmp_path):
    from tinydiff.conditional_diffusion import NoiseModel
    _roundtrip(tmp_path, NoiseModel(), init_state_dict("conditional_diffusion"), "")


def test_compiled_module_prefix_is_stripped(tmp_path):
    from tinydiff.conditional_diffusion_laion import NoiseModel
    sd = init_state_dict("conditional_diffusion_laion")
    _roundtrip(tmp_path, NoiseModel(time_dim=768), sd, "_orig_mod.")
    m = NoiseModel(time_dim=768)
    m.load_state_dict({"_orig_mod." + k: v for k, v in sd.items()})   # nn.Module API, same handling
    assert torch.equal(m.state_dict()["final_conv.weight"], sd["final_conv.weight"])


def test_dense_models_accept_prefix(tmp_path):
    from tinydiff.latent_diffusion import NoiseModel as Mlp
    from tinydiff.diffusion_transformer import NoiseModel as Dit
    _roundtrip(tmp_path, Mlp(), init_state_dict("latent_diffusion"), "_orig_mod.")
    _roundtrip(tmp_path, Dit(), init_state_dict("diffusion_transformer"), "")
