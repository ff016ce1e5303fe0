"""Host logic of seg_b200.optim: the factory the launcher installs as torch.optim.SGD leaves non-CUDA / non-plain uses
to the stock optimiser (the reference's CPU plumbing config keeps working) and the fused class is a torch.optim.SGD."""
import torch

from seg_b200 import optim


def test_factory_leaves_cpu_models_to_stock_sgd():
    ps = [torch.nn.Parameter(torch.zeros(3)), torch.nn.Parameter(torch.zeros(2))]
    groups = [{"params": filter(lambda p: p.requires_grad, ps[:1])}, {"params": iter(ps[1:]), "lr": 0.001}]  # as base_trainer.py:48-54
    o = optim.make_sgd(groups, lr=0.01, momentum=0.9, weight_decay=1e-4)
    assert type(o) is optim._STOCK_SGD and [len(g["params"]) for g in o.param_groups] == [1, 1]
    ps[0].grad, ps[1].grad = torch.ones(3), torch.ones(2)
    o.step()
    assert torch.allclose(ps[0], torch.full((3,), -0.01)) and torch.allclose(ps[1], torch.full((2,), -0.001))
    o2 = optim.make_sgd(iter(ps), lr=0.1)
    assert type(o2) is optim._STOCK_SGD


def test_fused_class_is_a_torch_sgd_with_the_same_state_layout():
    assert issubclass(optim.SGD, optim._STOCK_SGD)
    ps = [torch.nn.Parameter(torch.zeros(3))]
    o = optim.SGD(ps, lr=0.01, momentum=0.9, weight_decay=1e-4)
    ref = optim._STOCK_SGD(ps, lr=0.01, momentum=0.9, weight_decay=1e-4)
    keys = lambda opt: {k for k in opt.state_dict()["param_groups"][0]}  # noqa: E731
    assert keys(o) == keys(ref)
    import pytest
    with pytest.raises(NotImplementedError):
        optim.SGD(ps, lr=0.01, momentum=0.9, nesterov=True)
