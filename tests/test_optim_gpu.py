"""seg_b200.optim.SGD (one multi-tensor kernel per parameter group) against torch.optim.SGD on the same tensors:
two learning-rate groups, momentum, weight decay, a momentum / LR change mid-run (what OneCycle does), state_dict
round trip — the configuration base/base_trainer.py:46-57 builds."""
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from seg_b200.optim import SGD


def make(seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(64, 3, 7, 7), (64,), (256, 64, 1, 1), (19, 256, 1, 1), (19,), (1,), (2048, 512, 3, 3)]
    return [torch.nn.Parameter(torch.randn(s, generator=g).cuda()) for s in shapes]


def test_fused_sgd_matches_torch_sgd():
    a, b = make(1), make(1)
    kw = dict(lr=0.01, momentum=0.9, weight_decay=1e-4)
    ref = torch.optim.SGD([{"params": a[:4]}, {"params": a[4:], "lr": 0.001}], **kw)
    got = SGD([{"params": b[:4]}, {"params": b[4:], "lr": 0.001}], **kw)
    g = torch.Generator().manual_seed(2)
    for it in range(6):
        if it == 3:  # a schedule moves the rates and the momentum between steps
            for o in (ref, got):
                o.param_groups[0]["lr"], o.param_groups[1]["lr"] = 0.004, 0.0004
                o.param_groups[0]["momentum"] = o.param_groups[1]["momentum"] = 0.87
        for pa, pb in zip(a, b):
            gr = torch.randn(pa.shape, generator=g).cuda()
            pa.grad, pb.grad = gr.clone(), gr.clone()  # fresh gradient tensors every step, as autograd produces them
        if it == 4:
            a[5].grad = b[5].grad = None  # a parameter without a gradient is skipped
        ref.step()
        got.step()
        for i, (pa, pb) in enumerate(zip(a, b)):
            assert torch.allclose(pa, pb, rtol=1e-6, atol=2e-7), (it, i, (pa - pb).abs().max().item())
    sd = got.state_dict()
    assert set(sd["state"][0].keys()) == {"momentum_buffer"} and len(sd["state"]) == len(b)
    for i, (pa, pb) in enumerate(zip(a, b)):
        # buf = momentum*buf + d is one fma here and two rounded operations in torch: a few ulps of O(3) values
        assert torch.allclose(ref.state[pa]["momentum_buffer"], got.state[pb]["momentum_buffer"], rtol=1e-5, atol=2e-6), i
    fresh = SGD([{"params": b[:4]}, {"params": b[4:], "lr": 0.001}], **kw)
    fresh.load_state_dict(sd)
    ref2 = torch.optim.SGD([{"params": a[:4]}, {"params": a[4:], "lr": 0.001}], **kw)
    ref2.load_state_dict(ref.state_dict())
    for pa, pb in zip(a, b):
        gr = torch.randn(pa.shape, generator=g).cuda()
        pa.grad, pb.grad = gr.clone(), gr.clone()
    ref2.step()
    fresh.step()
    for i, (pa, pb) in enumerate(zip(a, b)):
        assert torch.allclose(pa, pb, rtol=1e-6, atol=2e-7), i


def test_refuses_cpu_parameters():
    p = torch.nn.Parameter(torch.zeros(3))
    p.grad = torch.ones(3)
    with pytest.raises(RuntimeError, match="no fallback"):
        SGD([p], lr=0.1, momentum=0.9).step()


def test_fused_sgd_step_invalidates_packed_weight_caches():
    """the kernel updates parameters through raw pointers; the eager plugin path caches packed bf16
    weights keyed on `param._version`, so SGD.step must bump the versions — otherwise every forward after step 1 keeps
    using the initial weights.  Two identical models, one stepped by torch.optim.SGD, one by seg_b200.optim.SGD: after
    each step the logits must have moved and must agree between the two."""
    import seg_b200
    from oracle import synth, weights

    sd = weights.deeplab_resnet_state_dict(19, "resnet14", seed=5)
    x, y = synth.make_batch(2, 65, 65, 19, 255, seed=77)
    x, y = x.cuda(), y.cuda()
    crit = seg_b200.CrossEntropyLoss2d(ignore_index=255)
    models, opts = [], []
    for cls in (torch.optim.SGD, SGD):
        m = seg_b200.DeepLab(19, backbone="resnet14", pretrained=False)
        m.load_state_dict(sd, strict=True)
        m.engine_dropout = False
        m = m.cuda().train()
        m.freeze_bn()  # a fixed function of the weights: differences are the optimiser's, not batch-statistics noise
        models.append(m)
        opts.append(cls([{"params": list(m.get_decoder_params())}, {"params": list(m.get_backbone_params()), "lr": 1e-4}],
                        lr=1e-3, momentum=0.9, weight_decay=1e-4))
    prev = None
    for it in range(3):
        outs = []
        for m, o in zip(models, opts):
            o.zero_grad(set_to_none=True)
            out = m(x)
            crit(out, y).backward()
            o.step()
            outs.append(out.detach().clone())
        scale = outs[0].abs().max().item()
        assert (outs[0] - outs[1]).abs().max().item() < 2e-2 * scale, it
        if prev is not None:  # the step changed what the forward computes
            assert (outs[1] - prev).abs().max().item() > 1e-4 * scale, "logits did not move: stale packed weights"
        prev = outs[1]
    for p in models[1].parameters():
        assert p._version >= 3
