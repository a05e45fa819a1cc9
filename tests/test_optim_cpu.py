"""Host logic of texture_gs_b200.optim.TextureAdam that needs no GPU."""
import pytest
import torch

from texture_gs_b200.optim import TextureAdam, _padded_storage


def test_padded_gradient_view_detection():
    buf = torch.zeros(6, 8, 8, 4)
    assert _padded_storage(buf[..., :3]) == buf.data_ptr()
    assert _padded_storage(torch.zeros(6, 8, 8, 5)[..., :3]) is None          # wrong pitch
    assert _padded_storage(buf[:, ::2, :, :3]) is None                         # not dense in the outer dims
    assert _padded_storage(buf[..., 1:4]) is not None and _padded_storage(buf[..., 1:4]) != buf.data_ptr()


def test_constructor_mirrors_torch_adam_and_rejects_what_the_reference_never_uses():
    p = torch.nn.Parameter(torch.zeros(6, 4, 4, 3))
    opt = TextureAdam([{"params": [p], "lr": 0.0025}], lr=0.0, eps=1e-15)      # models/texture_gaussian3d.py:139-143
    ref = torch.optim.Adam([{"params": [p], "lr": 0.0025}], lr=0.0, eps=1e-15)
    for k in ("lr", "betas", "eps", "weight_decay", "amsgrad", "maximize"):
        assert opt.param_groups[0][k] == ref.param_groups[0][k]
    assert set(opt.state_dict().keys()) == set(ref.state_dict().keys())
    with pytest.raises(NotImplementedError):
        TextureAdam([p], amsgrad=True)
    with pytest.raises(NotImplementedError):
        TextureAdam([p], weight_decay=0.1)


def test_step_on_cpu_tensors_raises_instead_of_falling_back():
    p = torch.nn.Parameter(torch.zeros(6, 4, 4, 3))
    p.grad = torch.ones_like(p)
    with pytest.raises(RuntimeError):
        TextureAdam([p], lr=0.1).step()
