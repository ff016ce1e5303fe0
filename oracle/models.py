"""Functional CPU restatement of the reference's hot-path models (TEST INFRASTRUCTURE — see oracle/__init__.py).

Every function takes the reference-named state_dict `sd` (tensors; set requires_grad on them to get gradients via
autograd) and follows the reference line by line:
  deeplab_forward   -> models/deeplabv3_plus.py:356-362 (DeepLab.forward), :55-63 (ResNet.forward),
                       :286-297 (ASSP.forward), :323-330 (Decoder.forward); torchvision Bottleneck (v1.5)
  pspnet_forward    -> models/pspnet.py:77-94, :31-38 (_PSPModule.forward); models/resnet.py:100-121,190-210
BatchNorm: F.batch_norm(train) == nn.BatchNorm2d in training mode (biased var to normalise, unbiased var into
running_var, momentum 0.1, eps 1e-5); running stats in `sd` are updated IN PLACE like the reference's buffers.
Dropout is the identity unless `dropout=True` (train-mode parity is checked with p=0).
"""
import torch
import torch.nn.functional as F

from .weights import RESNET_LAYERS

EPS = 1e-5
MOM = 0.1


def _bn(sd, name, x, train, sync_stats=None):
    w, b = sd[name + ".weight"], sd[name + ".bias"]
    rm, rv = sd[name + ".running_mean"], sd[name + ".running_var"]
    return F.batch_norm(x, rm, rv, w, b, train, MOM, EPS)


def _conv(sd, name, x, stride=1, pad=0, dil=1):
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride, pad, dil)


def _bottleneck(sd, p, x, stride, dil, train):
    """torchvision Bottleneck v1.5 / models/resnet.py:100-121: 1x1 -> 3x3(stride, dilation) -> 1x1, + identity."""
    out = F.relu(_bn(sd, p + "bn1", _conv(sd, p + "conv1", x), train))
    out = F.relu(_bn(sd, p + "bn2", _conv(sd, p + "conv2", out, stride, dil, dil), train))
    out = _bn(sd, p + "bn3", _conv(sd, p + "conv3", out), train)
    if (p + "downsample.0.weight") in sd:
        x = _bn(sd, p + "downsample.1", _conv(sd, p + "downsample.0", x, stride), train)
    return F.relu(out + x)


def deeplab_resnet_trunk(sd, x, backbone="resnet101", output_stride=16, train=True):
    """deeplabv3_plus.py:15-63: torchvision trunk with layer3/4 conv2 + downsample.0 stride->dilation surgery."""
    if output_stride == 16:
        s3, s4, d3, d4 = 2, 1, 1, 2
    else:
        s3, s4, d3, d4 = 1, 1, 2, 4
    x = _conv(sd, "backbone.layer0.0", x, 2, 3)
    x = F.relu(_bn(sd, "backbone.layer0.1", x, train))
    x = F.max_pool2d(x, 3, 2, 1)
    layers = RESNET_LAYERS[backbone]
    low = None
    # (layer index, first-block stride, dilation applied to EVERY conv2 of the layer)
    cfg = {1: (1, 1), 2: (2, 1), 3: (s3, d3 if output_stride == 8 else 1), 4: (s4, d4)}
    for li in (1, 2, 3, 4):
        stride, dil = cfg[li]
        for b in range(layers[li - 1]):
            x = _bottleneck(sd, f"backbone.layer{li}.{b}.", x, stride if b == 0 else 1, dil, train)
        if li == 1:
            low = x
    return x, low


def _sepconv(sd, name, x, stride, dil, train):
    """SeparableConv2d.forward (deeplabv3_plus.py:82-86): depthwise 3x3 ("same" padding = dilation) -> BN -> pointwise."""
    C = x.shape[1]
    y = F.conv2d(x, sd[name + ".conv1.weight"], None, stride, dil, dil, groups=C)
    y = _bn(sd, name + ".bn", y, train)
    return F.conv2d(y, sd[name + ".pointwise.weight"])


def _xblock(sd, p, x, stride, dil, train, exit_flow=False, use_1st_relu=True):
    """Block.forward (deeplabv3_plus.py:123-132).  rep[0] is an IN-PLACE ReLU on the block input, so the skip branch
    (and an identity skip) sees relu(x) — except in block1, built with use_1st_relu=False."""
    if use_1st_relu:
        x = F.relu(x)
    idx = 0 if not use_1st_relu else 1
    h = x
    for u in range(3):
        if u > 0:
            h = F.relu(h)
        s = stride if u == 2 else 1
        h = _sepconv(sd, f"{p}rep.{idx}", h, s, dil, train)
        h = _bn(sd, f"{p}rep.{idx + 1}", h, train)
        idx += 3
    if (p + "skip.weight") in sd:
        skip = _bn(sd, p + "skipbn", F.conv2d(x, sd[p + "skip.weight"], None, stride), train)
    else:
        skip = x
    return h + skip


def deeplab_xception_trunk(sd, x, output_stride=16, train=True):
    """Xception.forward (deeplabv3_plus.py:201-247): low_level_features = block1 output BEFORE the ReLU."""
    if output_stride == 16:
        b3_s, mf_d, ef_d = 2, 1, (1, 2)
    else:
        b3_s, mf_d, ef_d = 1, 2, (2, 4)
    x = F.relu(_bn(sd, "backbone.bn1", F.conv2d(x, sd["backbone.conv1.weight"], None, 2, 1), train))
    x = _bn(sd, "backbone.bn2", F.conv2d(x, sd["backbone.conv2.weight"], None, 1, 1), train)
    x = _xblock(sd, "backbone.block1.", x, 2, 1, train, use_1st_relu=False)
    low = x
    x = F.relu(x)
    x = _xblock(sd, "backbone.block2.", x, 2, 1, train)
    x = _xblock(sd, "backbone.block3.", x, b3_s, 1, train)
    for i in range(16):
        x = _xblock(sd, f"backbone.block{i + 4}.", x, 1, mf_d, train)
    x = _xblock(sd, "backbone.block20.", x, 1, ef_d[0], train, exit_flow=True)
    x = F.relu(x)
    for n in (3, 4, 5):
        x = F.relu(_bn(sd, f"backbone.bn{n}", _sepconv(sd, f"backbone.conv{n}", x, 1, ef_d[1], train), train))
    return x, low


def aspp(sd, x, output_stride=16, train=True, dropout=False):
    """deeplabv3_plus.py:286-297."""
    d = (1, 6, 12, 18) if output_stride == 16 else (1, 12, 24, 36)
    outs = []
    for i, k in zip((1, 2, 3, 4), (1, 3, 3, 3)):
        pad = 0 if k == 1 else d[i - 1]
        y = _conv(sd, f"ASSP.aspp{i}.0", x, 1, pad, d[i - 1])
        outs.append(F.relu(_bn(sd, f"ASSP.aspp{i}.1", y, train)))
    g = F.adaptive_avg_pool2d(x, 1)
    g = F.relu(_bn(sd, "ASSP.avg_pool.2", _conv(sd, "ASSP.avg_pool.1", g), train))
    outs.append(F.interpolate(g, size=x.shape[2:], mode="bilinear", align_corners=True))
    y = _conv(sd, "ASSP.conv1", torch.cat(outs, 1))
    y = F.relu(_bn(sd, "ASSP.bn1", y, train))
    return F.dropout(y, 0.5, True) if (dropout and train) else y


def decoder(sd, x, low, train=True, dropout=False):
    """deeplabv3_plus.py:323-330 (concat order: low-level first)."""
    low = F.relu(_bn(sd, "decoder.bn1", _conv(sd, "decoder.conv1", low), train))
    x = F.interpolate(x, size=low.shape[2:], mode="bilinear", align_corners=True)
    y = torch.cat((low, x), 1)
    y = F.relu(_bn(sd, "decoder.output.1", _conv(sd, "decoder.output.0", y, 1, 1), train))
    y = F.relu(_bn(sd, "decoder.output.4", _conv(sd, "decoder.output.3", y, 1, 1), train))
    if dropout and train:
        y = F.dropout(y, 0.1, True)
    return _conv(sd, "decoder.output.7", y)


def deeplab_forward(sd, x, backbone="resnet101", output_stride=16, train=True, dropout=False, return_lowres=False):
    """deeplabv3_plus.py:356-362.  Returns fp32 logits [B, C, H, W] (and the stride-4 logits if asked)."""
    H, W = x.shape[2:]
    if backbone == "xception":
        f, low = deeplab_xception_trunk(sd, x, output_stride, train)
    else:
        f, low = deeplab_resnet_trunk(sd, x, backbone, output_stride, train)
    f = aspp(sd, f, output_stride, train, dropout)
    lo = decoder(sd, f, low, train, dropout)
    out = F.interpolate(lo, size=(H, W), mode="bilinear", align_corners=True)
    return (out, lo) if return_lowres else out


def pspnet_forward(sd, x, backbone="resnet50", train=True, use_aux=True, dropout=False):
    """pspnet.py:77-94 over the deep-stem dilated ResNet (resnet.py:136-163,190-210).  Training returns (out, aux)."""
    size = x.shape[2:]
    x = F.relu(_bn(sd, "initial.0.1", _conv(sd, "initial.0.0", x, 2, 1), train))
    x = F.relu(_bn(sd, "initial.0.4", _conv(sd, "initial.0.3", x, 1, 1), train))
    x = F.relu(_bn(sd, "initial.1", _conv(sd, "initial.0.6", x, 1, 1), train))
    x = F.max_pool2d(x, 3, 2, 1)
    layers = RESNET_LAYERS[backbone]
    # (first-block stride, first-block dilation, other-block dilation): resnet.py:190-210
    cfg = {1: (1, 1, 1), 2: (2, 1, 1), 3: (1, 1, 2), 4: (1, 2, 4)}
    x_aux = None
    for li in (1, 2, 3, 4):
        stride, d0, d = cfg[li]
        for b in range(layers[li - 1]):
            x = _bottleneck(sd, f"layer{li}.{b}.", x, stride if b == 0 else 1, d0 if b == 0 else d, train)
        if li == 3:
            x_aux = x
    h, w = x.shape[2:]
    pyr = [x]
    for i, bins in enumerate((1, 2, 3, 6)):
        p = F.adaptive_avg_pool2d(x, bins)
        p = F.relu(_bn(sd, f"master_branch.0.stages.{i}.2", _conv(sd, f"master_branch.0.stages.{i}.1", p), train))
        pyr.append(F.interpolate(p, size=(h, w), mode="bilinear", align_corners=True))
    y = _conv(sd, "master_branch.0.bottleneck.0", torch.cat(pyr, 1), 1, 1)
    y = F.relu(_bn(sd, "master_branch.0.bottleneck.1", y, train))
    if dropout and train:
        y = F.dropout2d(y, 0.1, True)
    out = F.interpolate(_conv(sd, "master_branch.1", y), size=size, mode="bilinear", align_corners=False)
    if train and use_aux:
        a = F.relu(_bn(sd, "auxiliary_branch.1", _conv(sd, "auxiliary_branch.0", x_aux, 1, 1), train))
        if dropout:
            a = F.dropout2d(a, 0.1, True)
        aux = F.interpolate(_conv(sd, "auxiliary_branch.4", a), size=size, mode="bilinear", align_corners=False)
        return out, aux
    return out


def upernet_forward(sd, x, backbone="resnet101", train=True, dropout=False):
    """upernet.py:136-144 (UperNet.forward) with ResNet.forward :79-87 (output_stride 16: layer3 stride 2, layer4 every
    conv2 d=2), PSPModule.forward :31-38 (bins 1,2,4,6), FPN_fuse.forward :103-117 and up_and_add :89-90.
    `smooth_conv` is ONE conv applied three times (shared weights); the top-down path is NOT cumulative."""
    size = x.shape[2:]
    y = _conv(sd, "backbone.initial.0", x, 2, 3)
    y = F.relu(_bn(sd, "backbone.initial.1", y, train))
    y = F.max_pool2d(y, 3, 2, 1)
    layers = RESNET_LAYERS[backbone]
    cfg = {1: (1, 1), 2: (2, 1), 3: (2, 1), 4: (1, 2)}
    feats = []
    for li in (1, 2, 3, 4):
        stride, dil = cfg[li]
        for b in range(layers[li - 1]):
            y = _bottleneck(sd, f"backbone.layer{li}.{b}.", y, stride if b == 0 else 1, dil, train)
        feats.append(y)
    f4 = feats[-1]
    h, w = f4.shape[2:]
    pyr = [f4]
    for i, bins in enumerate((1, 2, 4, 6)):
        p = F.adaptive_avg_pool2d(f4, bins)
        p = F.relu(_bn(sd, f"PPN.stages.{i}.2", _conv(sd, f"PPN.stages.{i}.1", p), train))
        pyr.append(F.interpolate(p, size=(h, w), mode="bilinear", align_corners=True))
    y = F.relu(_bn(sd, "PPN.bottleneck.1", _conv(sd, "PPN.bottleneck.0", torch.cat(pyr, 1), 1, 1), train))
    if dropout and train:
        y = F.dropout2d(y, 0.1, True)
    feats[-1] = y
    # FPN_fuse: laterals (with bias) on features[1:], then up(f_i) + f_{i-1} from the ORIGINAL laterals
    f = [feats[0]] + [_conv(sd, f"FPN.conv1x1.{i}", feats[i + 1]) for i in range(3)]
    P = []
    for i in (3, 2, 1):
        up = F.interpolate(f[i], size=f[i - 1].shape[2:], mode="bilinear", align_corners=True) + f[i - 1]
        P.append(_conv(sd, "FPN.smooth_conv.0", up, 1, 1))
    P = list(reversed(P))
    P.append(f[-1])
    H, W = P[0].shape[2:]
    P[1:] = [F.interpolate(q, size=(H, W), mode="bilinear", align_corners=True) for q in P[1:]]
    y = F.relu(_bn(sd, "FPN.conv_fusion.1", _conv(sd, "FPN.conv_fusion.0", torch.cat(P, 1), 1, 1), train))
    y = _conv(sd, "head", y, 1, 1)
    return F.interpolate(y, size=size, mode="bilinear", align_corners=False)


def param_names(sd):
    """Names of trainable tensors (everything except BN running statistics / counters)."""
    return [k for k in sd if not (k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"))]


def clone_sd(sd, requires_grad=False):
    """Deep copy that preserves aliasing (UperNet's three smooth_conv entries are one tensor)."""
    out, seen = {}, {}
    names = set(param_names(sd))
    for k, v in sd.items():
        key = id(v)
        if key not in seen:
            t = v.clone()
            if requires_grad and k in names:
                t.requires_grad_(True)
            seen[key] = t
        out[k] = seen[key]
    return out
