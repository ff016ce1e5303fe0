"""Pins the CPU oracle (oracle/) against outputs of the UNMODIFIED reference captured by oracle/make_golden.py
(the reference ships no tests of its own — SURVEY.md §4 — so its own outputs are the only golden vectors)."""
import os

import numpy as np
import pytest
import torch

from oracle import losses as ol
from oracle import models as om
from oracle import synth, weights

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 2e-4  # oneDNN kernels differ between host CPUs; the generating machine reproduces these to ~1e-6


def close(a, b, rtol=RTOL):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.abs(b).max() + 1e-12
    err = np.abs(a - b).max() / scale
    assert err <= rtol, f"rel err {err:.3e} > {rtol:.1e}"


CASES = [
    ("deeplab_r101_65.npz", "deeplab", dict(num_classes=19, backbone="resnet101", seed=0), dict(output_stride=16), 9001),
    ("deeplab_r50_os8_65.npz", "deeplab", dict(num_classes=19, backbone="resnet50", seed=2), dict(output_stride=8), 9001),
    ("pspnet_r50_65.npz", "pspnet", dict(num_classes=21, backbone="resnet50", seed=1), dict(), 9002),
    ("deeplab_xception_65.npz", "deeplab", dict(num_classes=19, backbone="xception", seed=4), dict(output_stride=16), 9001),
    ("upernet_r50_64.npz", "upernet", dict(num_classes=150, backbone="resnet50", seed=3), dict(), 9004),
]


@pytest.mark.parametrize("fname,kind,wargs,fargs,xseed", CASES, ids=[c[0] for c in CASES])
def test_model_oracle_matches_reference_golden(fname, kind, wargs, fargs, xseed):
    g = np.load(os.path.join(GOLD, fname))
    nc = wargs["num_classes"]
    ignore = 255
    if kind == "deeplab" and wargs["backbone"] == "xception":
        sd = weights.deeplab_xception_state_dict(nc, seed=wargs["seed"], randomize_bn=True)
    elif kind == "deeplab":
        sd = weights.deeplab_resnet_state_dict(nc, wargs["backbone"], seed=wargs["seed"], randomize_bn=True)
    elif kind == "upernet":
        sd = weights.upernet_state_dict(nc, wargs["backbone"], seed=wargs["seed"], randomize_bn=True)
        ignore = -1
    else:
        sd = weights.pspnet_state_dict(nc, wargs["backbone"], seed=wargs["seed"], randomize_bn=True)
    size = 64 if kind == "upernet" else 65
    x, y = synth.make_batch(2, size, size, nc, ignore, seed=xseed)
    sd = om.clone_sd(sd, requires_grad=True)
    if kind == "deeplab":
        out = om.deeplab_forward(sd, x, backbone=wargs["backbone"], train=True, **fargs)
        loss = ol.cross_entropy2d(out, y, 255)
        aux = None
    elif kind == "upernet":
        out = om.upernet_forward(sd, x, backbone=wargs["backbone"], train=True)
        loss = ol.cross_entropy2d(out, y, -1)
        aux = None
    else:
        out, aux = om.pspnet_forward(sd, x, backbone=wargs["backbone"], train=True)
        loss = ol.cross_entropy2d(out, y, 255) + 0.4 * ol.cross_entropy2d(aux, y, 255)
    loss.backward()
    close(out.detach()[:, :, ::3, ::3].numpy(), g["logits_sub"])
    close(out.detach().double().sum((2, 3)).numpy(), g["logits_sum"])
    assert (out.detach().argmax(1).numpy() == g["argmax"]).mean() > 0.9995
    close(loss.item(), g["loss"], 1e-5)
    if aux is not None:
        close(aux.detach()[:, :, ::3, ::3].numpy(), g["aux_sub"])
    names = [str(n) for n in g["param_names"]]
    uniq, seen = [], set()
    for n in om.param_names(sd):  # named_parameters() lists a shared tensor once (UperNet's smooth conv)
        if id(sd[n]) not in seen:
            seen.add(id(sd[n]))
            uniq.append(n)
    assert names == uniq, "oracle parameter order/names differ from the reference's named_parameters()"
    if g["grad_norms"].size:  # (the reference cannot run backward for Xception on torch >= 1.5; see make_golden.py)
        gn = np.array([sd[n].grad.double().norm().item() for n in names])
        close(gn, g["grad_norms"], 2e-3)
    else:
        assert "inplace" in str(g["backward_error"])
    for k in g.files:
        if k.startswith("grad/"):
            close(sd[k[5:]].grad.numpy(), g[k], 2e-3)
        elif k.startswith("rm/"):
            close(sd[k[3:] + ".running_mean"].numpy(), g[k])
        elif k.startswith("rv/"):
            close(sd[k[3:] + ".running_var"].numpy(), g[k])
    with torch.no_grad():
        if kind == "deeplab":
            ev = om.deeplab_forward(sd, x, backbone=wargs["backbone"], train=False, **fargs)
        elif kind == "upernet":
            ev = om.upernet_forward(sd, x, backbone=wargs["backbone"], train=False)
        else:
            ev = om.pspnet_forward(sd, x, backbone=wargs["backbone"], train=False)
    close(ev.double().sum((2, 3)).numpy(), g["eval_logits_sum"])


@pytest.mark.parametrize("tag,ignore", [("c7", 255), ("c150", -1)])
def test_loss_oracle_matches_reference_golden(tag, ignore):
    g = np.load(os.path.join(GOLD, "losses_syncbn.npz"))
    fns = {"ce": ol.cross_entropy2d, "dice": ol.dice_loss, "lovasz": ol.lovasz_softmax, "ce_dice": ol.ce_dice_loss}
    for name, fn in fns.items():
        key = f"{tag}/{name}/loss"
        if key not in g.files:
            continue
        lg = torch.from_numpy(g[f"{tag}/logits"]).clone().requires_grad_(True)
        tg = torch.from_numpy(g[f"{tag}/{name}/target_in"] if f"{tag}/{name}/target_in" in g.files else g[f"{tag}/target"]).clone()
        loss = fn(lg, tg, ignore_index=ignore) if name != "dice" else fn(lg, tg, 1.0, ignore)
        loss.backward()
        close(loss.item(), g[key], 1e-5)
        close(lg.grad.numpy(), g[f"{tag}/{name}/grad"], 1e-4)
        assert (tg.numpy() == g[f"{tag}/{name}/target_after"]).all(), "target mutation (losses.py:40-42) not reproduced"


def test_syncbn_formula_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "losses_syncbn.npz"))
    x = torch.from_numpy(g["syncbn/x"])
    flat = x.permute(1, 0, 2, 3).reshape(x.shape[1], -1)
    mean, istd, unbias = ol.syncbn_mean_istd(flat.sum(1), (flat * flat).sum(1), flat.shape[1])
    close(mean.numpy(), g["syncbn/mean"], 1e-6)
    close(istd.numpy(), g["syncbn/inv_std"], 1e-6)
    close((0.9 * torch.ones(8) + 0.1 * unbias).numpy(), g["syncbn/running_var"], 1e-6)
    # clamp(var, eps) differs from (var + eps) where var is tiny — the quirk the engine must reproduce in sync mode
    var = flat.var(1, unbiased=False)
    assert (var < 1e-5).any()


def test_state_dict_inventory():
    assert len(weights.deeplab_resnet_state_dict(19, "resnet101")) == 680
    assert len(weights.pspnet_state_dict(21, "resnet50")) == 370
    assert len(weights.deeplab_xception_state_dict(19)) == 848
    usd = weights.upernet_state_dict(150, "resnet101")
    assert sum(v.numel() for k, v in {id(v): v for k, v in usd.items() if k in set(om.param_names(usd))}.items()) == 126414038
    sd = weights.pspnet_state_dict(19, "resnet50")
    names = set(om.param_names(sd))
    n = sum(v.numel() for k, v in sd.items() if k in names)
    assert n == 51444710  # tutorial.ipynb:5408, the reference's one published golden scalar


def test_eval_metrics_oracle_matches_reference_golden():
    """oracle/metrics.py (restatement of utils/metrics.py:41-67) against values produced by the reference itself."""
    from oracle import metrics as om
    g = np.load(os.path.join(GOLD, "metrics.npz"))
    for tag, K in (("c7", 7), ("c19", 19), ("c150", 150)):
        out = om.eval_metrics(g[f"{tag}/logits"], g[f"{tag}/target"], K)
        assert int(out[0]) == int(g[f"{tag}/correct"]) and int(out[1]) == int(g[f"{tag}/labeled"]), tag
        assert np.array_equal(out[2], g[f"{tag}/inter"]) and np.array_equal(out[3], g[f"{tag}/union"]), tag
        assert int(out[0]) > 0 and out[2].sum() == out[0]  # the vectors exercise the counters
