"""Host-side logic of the classifier-free-guidance extension; CPU only."""
import torch


def test_cfg_label_dropout_statistics_and_determinism():
    from tinydiff.conditional_diffusion import cfg_drop_labels
    y = torch.randint(0, 10, (20000,), generator=torch.Generator().manual_seed(0))
    a = cfg_drop_labels(y, 0.1, 10, generator=torch.Generator().manual_seed(1))
    b = cfg_drop_labels(y, 0.1, 10, generator=torch.Generator().manual_seed(1))
    assert torch.equal(a, b)
    dropped = a == 10
    assert 0.08 < dropped.float().mean() < 0.12
    assert torch.equal(a[~dropped], y[~dropped])
    assert torch.equal(cfg_drop_labels(y, 0.0, 10), y)
    assert bool((cfg_drop_labels(y, 1.0, 10) == 10).all())


def test_null_label_row_is_part_of_the_state_dict():
    from tinydiff.conditional_diffusion import NoiseModel
    m = NoiseModel(num_classes=11)
    assert m.state_dict()["class_embedding.weight"].shape == (11, 256)
