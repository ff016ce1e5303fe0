"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference) on oracle-made
weights and synthetic inputs.  Run in the build container only (the GPU box has no /root/reference):

    python -m oracle.make_golden

Shims are the non-invasive ones of SURVEY.md §8c (skimage stub; pretrained=False; dropout p=0 for train parity).
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")


def import_reference():
    sys.path.insert(0, REF)
    for n in ("skimage", "skimage.filters"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["skimage.filters"].gaussian = None
    import models  # noqa
    from utils import losses  # noqa
    return models, losses


def no_dropout(m):
    for mod in m.modules():
        if isinstance(mod, (torch.nn.Dropout, torch.nn.Dropout2d)):
            mod.p = 0.0


def import_upernet_shims():
    """models/upernet.py:133 references undefined names: define them as module globals, non-invasively."""
    import models.upernet as U
    from utils import helpers
    U.freeze_backbone = False
    U.set_trainable = helpers.set_trainable


SMALL_GRADS = {
    "deeplab": ["decoder.output.7.weight", "decoder.output.7.bias", "ASSP.bn1.weight", "ASSP.bn1.bias",
                "backbone.layer0.0.weight", "backbone.layer4.2.bn3.weight", "decoder.conv1.weight"],
    "pspnet": ["master_branch.1.weight", "master_branch.1.bias", "auxiliary_branch.4.bias", "initial.0.0.weight",
               "master_branch.0.stages.2.2.weight", "layer4.2.bn3.bias"],
    "xception": ["decoder.output.7.bias", "backbone.conv1.weight", "backbone.block1.rep.0.conv1.weight", "backbone.block20.skipbn.weight",
                 "backbone.conv5.pointwise.weight", "backbone.block10.rep.4.bn.bias"],
    "upernet": ["head.bias", "FPN.smooth_conv.0.bias", "FPN.conv1x1.2.bias", "FPN.conv_fusion.1.weight", "PPN.stages.3.2.bias",
                "backbone.initial.0.weight"],
}
BN_TRACK = {"deeplab": ["backbone.layer0.1", "ASSP.avg_pool.2", "decoder.output.4"],
            "pspnet": ["initial.1", "master_branch.0.stages.0.2", "auxiliary_branch.1"],
            "xception": ["backbone.bn1", "backbone.block1.rep.0.bn", "backbone.bn5"],
            "upernet": ["backbone.initial.1", "PPN.bottleneck.1", "FPN.conv_fusion.1"]}


def model_golden(kind, ref, sd, x, y, crit, fname):
    ref.load_state_dict(sd, strict=True)  # proves the oracle's key names and shapes are the reference's
    no_dropout(ref)
    ref.train()
    out = ref(x)
    if kind == "pspnet":  # (kind "xception" is a DeepLab: single output)
        loss = crit(out[0], y) + 0.4 * crit(out[1], y)  # trainer.py:60-61
        logits, aux = out
    else:
        loss = crit(out, y)
        logits, aux = out, None
    names = [n for n, _ in ref.named_parameters()]
    backward_error = ""
    try:
        loss.backward()
        gn = np.array([p.grad.double().norm().item() for _, p in ref.named_parameters()])
    except RuntimeError as e:
        # DeepLab/Xception: `x = F.relu(x)` followed by block2's in-place ReLU trips autograd's version check on
        # torch >= 1.5 (deeplabv3_plus.py:210,99-101) — the reference cannot back-propagate this model here; the
        # golden then pins the forward pass only and gradients are pinned through the oracle's equivalent math.
        backward_error = str(e).splitlines()[0][:200]
        gn = np.zeros(0)
    rec = {
        "param_names": np.array(names),
        "grad_norms": gn,
        "backward_error": np.array(backward_error),
        "loss": np.float64(loss.item()),
        "logits_sub": logits.detach()[:, :, ::3, ::3].numpy(),
        "logits_sum": logits.detach().double().sum((2, 3)).numpy(),
        "argmax": logits.detach().argmax(1).to(torch.uint8).numpy(),
    }
    if aux is not None:
        rec["aux_sub"] = aux.detach()[:, :, ::3, ::3].numpy()
        rec["aux_sum"] = aux.detach().double().sum((2, 3)).numpy()
    params = dict(ref.named_parameters())
    for n in SMALL_GRADS[kind]:
        if not backward_error:
            rec["grad/" + n] = params[n].grad.numpy()
    rs = ref.state_dict()
    for n in BN_TRACK[kind]:
        rec["rm/" + n] = rs[n + ".running_mean"].numpy()
        rec["rv/" + n] = rs[n + ".running_var"].numpy()
    # eval-mode forward with the updated running stats
    ref.eval()
    with torch.no_grad():
        ev = ref(x)
    rec["eval_logits_sum"] = ev.double().sum((2, 3)).numpy()
    np.savez_compressed(os.path.join(OUT, fname), **rec)
    print(fname, "loss", loss.item(), "params", len(names))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(8)
    models, losses = import_reference()
    sys.path.insert(0, ROOT)
    from oracle import synth, weights

    # ---- DeepLabV3+/ResNet-101 (config C3 architecture, reduced spatial size) ----
    sd = weights.deeplab_resnet_state_dict(19, "resnet101", seed=0, randomize_bn=True)
    x, y = synth.make_batch(2, 65, 65, 19, 255, seed=9001)
    ref = models.DeepLab(19, backbone="resnet101", pretrained=False)
    model_golden("deeplab", ref, sd, x, y, losses.CrossEntropyLoss2d(ignore_index=255), "deeplab_r101_65.npz")
    # output_stride 8 variant
    sd = weights.deeplab_resnet_state_dict(19, "resnet50", seed=2, randomize_bn=True)
    ref = models.DeepLab(19, backbone="resnet50", pretrained=False, output_stride=8)
    model_golden("deeplab", ref, sd, x, y, losses.CrossEntropyLoss2d(ignore_index=255), "deeplab_r50_os8_65.npz")

    # ---- DeepLabV3+/Aligned-Xception (config C4 architecture, reduced spatial size) ----
    sd = weights.deeplab_xception_state_dict(19, seed=4, randomize_bn=True)
    ref = models.DeepLab(19, backbone="xception", pretrained=False)
    model_golden("xception", ref, sd, x, y, losses.CrossEntropyLoss2d(ignore_index=255), "deeplab_xception_65.npz")

    # ---- PSPNet/ResNet-50 (config C2 architecture, reduced spatial size) ----
    sd = weights.pspnet_state_dict(21, "resnet50", seed=1, randomize_bn=True)
    x, y = synth.make_batch(2, 65, 65, 21, 255, seed=9002)
    ref = models.PSPNet(21, backbone="resnet50", pretrained=False)
    model_golden("pspnet", ref, sd, x, y, losses.CrossEntropyLoss2d(ignore_index=255), "pspnet_r50_65.npz")

    # ---- UperNet/ResNet-50 (config C5 architecture, reduced size; ADE20K-style labels -1..149, ignore_index -1) ----
    import_upernet_shims()
    sd = weights.upernet_state_dict(150, "resnet50", seed=3, randomize_bn=True)
    x, y = synth.make_batch(2, 64, 64, 150, -1, seed=9004)
    ref = models.UperNet(150, backbone="resnet50", pretrained=False)
    model_golden("upernet", ref, sd, x, y, losses.CrossEntropyLoss2d(ignore_index=-1), "upernet_r50_64.npz")

    # ---- losses (utils/losses.py) ----
    g = torch.Generator().manual_seed(9003)
    rec = {}
    for tag, C, ignore in (("c7", 7, 255), ("c150", 150, -1)):
        logits = (torch.randn(2, C, 12, 13, generator=g) * 2).requires_grad_(True)
        target = torch.randint(0, C, (2, 12, 13), generator=g)
        target[:, :2, :] = ignore
        rec[f"{tag}/logits"] = logits.detach().numpy()
        rec[f"{tag}/target"] = target.numpy()
        for name, cls in (("ce", losses.CrossEntropyLoss2d), ("dice", losses.DiceLoss), ("lovasz", losses.LovaszSoftmax),
                          ("ce_dice", losses.CE_DiceLoss)):
            if name in ("dice", "ce_dice") and ignore == -1:
                continue  # make_one_hot cannot scatter index -1; the reference never pairs Dice with ADE20K labels
            crit = cls(ignore_index=ignore)
            lg = logits.detach().clone().requires_grad_(True)
            tg = target.clone()
            if name == "ce_dice":
                # with ignored pixels DiceLoss mutates `target` after CE saved it -> autograd version error on
                # torch >= 1.5; the reference combination is only usable without ignored pixels
                tg[tg == ignore] = 0
                rec[f"{tag}/{name}/target_in"] = tg.numpy().copy()
            loss = crit(lg, tg)
            loss.backward()
            rec[f"{tag}/{name}/loss"] = np.float64(loss.item())
            rec[f"{tag}/{name}/grad"] = lg.grad.numpy()
            rec[f"{tag}/{name}/target_after"] = tg.numpy()  # DiceLoss mutates it (losses.py:40-42)
    # ---- SyncBN statistics formula (sync_batchnorm/batchnorm.py:128-145) ----
    from utils.sync_batchnorm.batchnorm import SynchronizedBatchNorm2d
    bn = SynchronizedBatchNorm2d(8)
    xs = torch.randn(4, 8, 5, 5, generator=g) * torch.tensor([1e-3, 0.1, 1, 2, 3, 1e-4, 5, 1]).view(1, 8, 1, 1)
    flat = xs.permute(1, 0, 2, 3).reshape(8, -1)
    mean, istd = bn._compute_mean_std(flat.sum(1), (flat * flat).sum(1), flat.shape[1])
    rec["syncbn/x"] = xs.numpy()
    rec["syncbn/mean"] = mean.numpy()
    rec["syncbn/inv_std"] = istd.numpy()
    rec["syncbn/running_mean"] = bn.running_mean.numpy()
    rec["syncbn/running_var"] = bn.running_var.numpy()
    np.savez_compressed(os.path.join(OUT, "losses_syncbn.npz"), **rec)
    print("losses_syncbn.npz written")


if __name__ == "__main__":
    main()
