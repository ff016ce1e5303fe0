"""Host-logic check of the engine on the CPU box: the kernel wrappers are swapped for the ATen emulation in
tests/cpu_emulation.py (test-only), and a full train step (forward, CE, backward) of the drop-in models is compared
with the oracle.  This validates tape ordering, gradient accumulation across branches, concat slices, the autograd
bridge and state_dict plumbing without a GPU; the kernels themselves are validated by the -m gpu tests."""
import pytest
import torch

import cpu_emulation as emu
from oracle import losses as ol
from oracle import models as om
from oracle import synth, weights


@pytest.fixture()
def emulated(monkeypatch):
    from seg_b200 import engine, nets
    from seg_b200 import losses as plosses
    for mod in (engine, nets, plosses):
        monkeypatch.setattr(mod, "ops", emu)
    # fp32 storage on both sides of the seam: this test is about host logic, not bf16 quantisation
    monkeypatch.setattr(engine, "ACT_DTYPE", torch.float32)
    monkeypatch.setattr(emu, "ACT_DTYPE", torch.float32)
    monkeypatch.setattr(nets._EngineModel, "_check_input", lambda self, x: None)
    return nets


def relerr(a, b):
    return ((a.detach().double() - b.detach().double()).abs().max() / (b.detach().double().abs().max() + 1e-12)).item()


@pytest.mark.parametrize("kind,backbone,kw", [("deeplab", "resnet50", dict(output_stride=16)), ("pspnet", "resnet50", dict()),
                                              ("upernet", "resnet50", dict()), ("deeplab", "xception", dict(output_stride=16))])
def test_train_step_host_logic(emulated, kind, backbone, kw):
    import seg_b200
    nc = 7
    if kind == "deeplab" and backbone == "xception":
        sd = weights.deeplab_xception_state_dict(nc, seed=5, randomize_bn=True, **kw)
        m = emulated.DeepLab(nc, backbone=backbone, **kw)
    elif kind == "deeplab":
        sd = weights.deeplab_resnet_state_dict(nc, backbone, seed=5, randomize_bn=True)
        m = emulated.DeepLab(nc, backbone=backbone, **kw)
    elif kind == "upernet":
        sd = weights.upernet_state_dict(nc, backbone, seed=5, randomize_bn=True)
        m = emulated.UperNet(nc, backbone=backbone, **kw)
    else:
        sd = weights.pspnet_state_dict(nc, backbone, seed=5, randomize_bn=True)
        m = emulated.PSPNet(nc, backbone=backbone, **kw)
    m.load_state_dict(sd, strict=True)
    m.engine_dropout = False
    m.train()
    size = 97 if backbone == "xception" else 49  # 5 stride-2 stages: keep BatchNorm sample counts sane
    x, y = synth.make_batch(2, size, size, nc, 255, seed=77)
    osd = om.clone_sd(sd, requires_grad=True)
    if kind == "deeplab":
        ref = om.deeplab_forward(osd, x, backbone=backbone, train=True, **kw)
        ref_loss = ol.cross_entropy2d(ref, y, 255)
    elif kind == "upernet":
        ref = om.upernet_forward(osd, x, backbone=backbone, train=True)
        ref_loss = ol.cross_entropy2d(ref, y, 255)
    else:
        ref, ref_aux = om.pspnet_forward(osd, x, backbone=backbone, train=True)
        ref_loss = ol.cross_entropy2d(ref, y, 255) + 0.4 * ol.cross_entropy2d(ref_aux, y, 255)
    ref_loss.backward()
    crit = seg_b200.CrossEntropyLoss2d(ignore_index=255)
    # CrossEntropyLoss2d refuses CPU tensors; call its autograd function directly (emulated kernels)
    from seg_b200.losses import _CEFn
    out = m(x)
    if kind == "pspnet":
        out, aux = out
        loss = _CEFn.apply(out, y, 255) + 0.4 * _CEFn.apply(aux, y, 255)
    else:
        loss = _CEFn.apply(out, y, 255)
    loss.backward()
    assert relerr(out, ref) < 2e-3
    assert abs(loss.item() - ref_loss.item()) < 1e-3 * abs(ref_loss.item())
    cos_min, worst = 1.0, None
    norms = torch.tensor([osd[n].grad.double().norm().item() for n, _ in m.named_parameters()])
    floor = 1e-4 * norms.median().item()
    for n, p in m.named_parameters():
        assert p.grad is not None, n
        assert osd[n].grad is not None, n
        if osd[n].grad.double().norm().item() < floor:
            # analytically zero gradient (e.g. the bias of SeparableConv2d.bn: pointwise conv + batch-stat BN downstream
            # are invariant to a per-channel shift) — both sides hold rounding noise only
            assert p.grad.double().norm().item() < 100 * floor, n
            continue
        c = torch.nn.functional.cosine_similarity(p.grad.double().flatten(), osd[n].grad.double().flatten(), dim=0).item()
        if c < cos_min:
            cos_min, worst = c, n
    # ReLU-mask flips on near-zero pre-activations are the only residual difference
    assert cos_min > 0.97, f"min cosine {cos_min} at {worst}"
    esd = m.state_dict()
    for k in esd:
        if k.endswith("running_mean") or k.endswith("running_var"):
            assert relerr(esd[k], osd[k]) < 2e-3, k
    # eval forward uses the running statistics and records nothing
    m.eval()
    with torch.no_grad():
        ev = m(x)
        ev_ref = {"deeplab": lambda: om.deeplab_forward(osd, x, backbone=backbone, train=False, **kw),
                  "pspnet": lambda: om.pspnet_forward(osd, x, backbone=backbone, train=False),
                  "upernet": lambda: om.upernet_forward(osd, x, backbone=backbone, train=False)}[kind]()
    assert isinstance(ev, torch.Tensor) and relerr(ev, ev_ref) < 2e-3


def test_cpu_input_is_refused():
    import seg_b200
    m = seg_b200.DeepLab(5, backbone="resnet50")
    with pytest.raises(RuntimeError, match="no CPU"):
        m(torch.zeros(1, 3, 33, 33))
    with pytest.raises(RuntimeError, match="no CPU"):
        seg_b200.CrossEntropyLoss2d()(torch.zeros(1, 5, 4, 4), torch.zeros(1, 4, 4, dtype=torch.long))
