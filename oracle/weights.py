"""Deterministic state_dict factories with the reference's parameter names and OIHW fp32 shapes.

numpy's RandomState is used (bit-stable across machines) instead of torch initialisers; distributions follow
utils/helpers.py:12-22 (kaiming-normal convs, BN gamma=1 / beta=1e-4) and torchvision's ResNet init, with an
option to randomise BN affine/running statistics so parity tests exercise them.
"""
import numpy as np
import torch


class _Gen:
    def __init__(self, seed, randomize_bn):
        self.rs = np.random.RandomState(seed)
        self.randomize_bn = randomize_bn
        self.sd = {}

    def conv(self, name, cout, cin, k, bias=False):
        fan_in = cin * k * k
        w = self.rs.standard_normal((cout, cin, k, k)).astype(np.float32) * np.float32(np.sqrt(2.0 / fan_in))
        self.sd[name + ".weight"] = torch.from_numpy(w)
        if bias:
            self.sd[name + ".bias"] = torch.from_numpy((self.rs.standard_normal(cout) * 0.1).astype(np.float32))

    def dwconv(self, name, c):
        """depthwise 3x3 weight [C,1,3,3]; kaiming-normal with fan_in = 9 (utils/helpers.py:15 on a grouped conv)."""
        w = self.rs.standard_normal((c, 1, 3, 3)).astype(np.float32) * np.float32(np.sqrt(2.0 / 9.0))
        self.sd[name + ".weight"] = torch.from_numpy(w)

    def sep(self, name, cin, cout):
        """SeparableConv2d (deeplabv3_plus.py:70-86): conv1 (depthwise) -> bn -> pointwise."""
        self.dwconv(name + ".conv1", cin)
        self.bn(name + ".bn", cin)
        self.conv(name + ".pointwise", cout, cin, 1)

    def bn(self, name, c, beta0=1e-4):
        if self.randomize_bn:
            self.sd[name + ".weight"] = torch.from_numpy((0.5 + self.rs.random_sample(c)).astype(np.float32))
            self.sd[name + ".bias"] = torch.from_numpy((self.rs.standard_normal(c) * 0.1).astype(np.float32))
            self.sd[name + ".running_mean"] = torch.from_numpy((self.rs.standard_normal(c) * 0.1).astype(np.float32))
            self.sd[name + ".running_var"] = torch.from_numpy((0.5 + self.rs.random_sample(c)).astype(np.float32))
        else:
            self.sd[name + ".weight"] = torch.ones(c)
            self.sd[name + ".bias"] = torch.full((c,), beta0)
            self.sd[name + ".running_mean"] = torch.zeros(c)
            self.sd[name + ".running_var"] = torch.ones(c)
        self.sd[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.int64)


RESNET_LAYERS = {"resnet50": (3, 4, 6, 3), "resnet101": (3, 4, 23, 3), "resnet152": (3, 8, 36, 3),
                 "resnet14": (1, 1, 1, 1)}  # resnet14: shallow test-only trunk (same Bottleneck code path)


def _tv_bottleneck_layers(g, prefix, layers, inplanes=64):
    """torchvision ResNet layer1..4 with Bottleneck blocks (expansion 4)."""
    planes_list = (64, 128, 256, 512)
    for li, (nblocks, planes) in enumerate(zip(layers, planes_list), start=1):
        for b in range(nblocks):
            p = f"{prefix}layer{li}.{b}."
            g.conv(p + "conv1", planes, inplanes, 1)
            g.bn(p + "bn1", planes, beta0=0.0)
            g.conv(p + "conv2", planes, planes, 3)
            g.bn(p + "bn2", planes, beta0=0.0)
            g.conv(p + "conv3", planes * 4, planes, 1)
            g.bn(p + "bn3", planes * 4, beta0=0.0)
            if b == 0:
                g.conv(p + "downsample.0", planes * 4, inplanes, 1)
                g.bn(p + "downsample.1", planes * 4, beta0=0.0)
            inplanes = planes * 4
    return inplanes


def deeplab_resnet_state_dict(num_classes, backbone="resnet101", seed=0, randomize_bn=False, in_channels=3):
    """Keys/shapes of models.DeepLab(num_classes, backbone='resnet*', pretrained=False).state_dict()
    (deeplabv3_plus.py:15-63,253-330): 680 entries for resnet101."""
    g = _Gen(seed, randomize_bn)
    g.conv("backbone.layer0.0", 64, in_channels, 7)
    g.bn("backbone.layer0.1", 64)
    _tv_bottleneck_layers(g, "backbone.", RESNET_LAYERS[backbone])
    for i, k in zip((1, 2, 3, 4), (1, 3, 3, 3)):
        g.conv(f"ASSP.aspp{i}.0", 256, 2048, k)
        g.bn(f"ASSP.aspp{i}.1", 256)
    g.conv("ASSP.avg_pool.1", 256, 2048, 1)
    g.bn("ASSP.avg_pool.2", 256)
    g.conv("ASSP.conv1", 256, 1280, 1)
    g.bn("ASSP.bn1", 256)
    g.conv("decoder.conv1", 48, 256, 1)
    g.bn("decoder.bn1", 48)
    g.conv("decoder.output.0", 256, 304, 3)
    g.bn("decoder.output.1", 256)
    g.conv("decoder.output.3", 256, 256, 3)
    g.bn("decoder.output.4", 256)
    g.conv("decoder.output.7", num_classes, 256, 1, bias=True)
    return g.sd


def pspnet_state_dict(num_classes, backbone="resnet50", seed=0, randomize_bn=False, in_channels=3, use_aux=True):
    """Keys/shapes of models.PSPNet(num_classes, backbone='resnet50', pretrained=False).state_dict()
    (pspnet.py:41-75 over the deep-stem ResNet of resnet.py:124-212): 370 entries for resnet50 with aux."""
    g = _Gen(seed, randomize_bn)
    # deep stem: initial = [conv1 (Sequential of 3 convs + 2 BN/ReLU), bn1, relu, maxpool]  (pspnet.py:49; resnet.py:136-151)
    g.conv("initial.0.0", 64, in_channels, 3)
    g.bn("initial.0.1", 64, beta0=0.0)
    g.conv("initial.0.3", 64, 64, 3)
    g.bn("initial.0.4", 64, beta0=0.0)
    g.conv("initial.0.6", 128, 64, 3)
    g.bn("initial.1", 128, beta0=0.0)
    planes_list = (64, 128, 256, 512)
    inplanes = 128
    for li, (nblocks, planes) in enumerate(zip(RESNET_LAYERS[backbone], planes_list), start=1):
        for b in range(nblocks):
            p = f"layer{li}.{b}."
            g.conv(p + "conv1", planes, inplanes, 1)
            g.bn(p + "bn1", planes, beta0=0.0)
            g.conv(p + "conv2", planes, planes, 3)
            g.bn(p + "bn2", planes, beta0=0.0)
            g.conv(p + "conv3", planes * 4, planes, 1)
            g.bn(p + "bn3", planes * 4, beta0=0.0)
            if b == 0:
                g.conv(p + "downsample.0", planes * 4, inplanes, 1)
                g.bn(p + "downsample.1", planes * 4, beta0=0.0)
            inplanes = planes * 4
    m_out = 2048
    for i in range(4):
        g.conv(f"master_branch.0.stages.{i}.1", m_out // 4, m_out, 1)
        g.bn(f"master_branch.0.stages.{i}.2", m_out // 4)
    g.conv("master_branch.0.bottleneck.0", m_out // 4, m_out * 2, 3)
    g.bn("master_branch.0.bottleneck.1", m_out // 4)
    g.conv("master_branch.1", num_classes, m_out // 4, 1, bias=True)
    if use_aux:
        g.conv("auxiliary_branch.0", m_out // 4, m_out // 2, 3)
        g.bn("auxiliary_branch.1", m_out // 4)
        g.conv("auxiliary_branch.4", num_classes, m_out // 4, 1, bias=True)
    return g.sd


def upernet_state_dict(num_classes, backbone="resnet101", seed=0, randomize_bn=False, in_channels=3, fpn_out=256):
    """Keys/shapes of models.UperNet(num_classes, backbone='resnet*', pretrained=False).state_dict()
    (upernet.py:40-87,9-38,92-134).  FPN.smooth_conv.{0,1,2} alias ONE Conv2d (upernet.py:98-99): the three
    state_dict entries share the same tensor."""
    g = _Gen(seed, randomize_bn)
    g.conv("backbone.initial.0", 64, in_channels, 7)
    g.bn("backbone.initial.1", 64)
    _tv_bottleneck_layers(g, "backbone.", RESNET_LAYERS[backbone])
    for i in range(4):
        g.conv(f"PPN.stages.{i}.1", 512, 2048, 1)
        g.bn(f"PPN.stages.{i}.2", 512, beta0=0.0)
    g.conv("PPN.bottleneck.0", 2048, 4096, 3)
    g.bn("PPN.bottleneck.1", 2048, beta0=0.0)
    for i, c in enumerate((512, 1024, 2048)):
        g.conv(f"FPN.conv1x1.{i}", fpn_out, c, 1, bias=True)
    g.conv("FPN.smooth_conv.0", fpn_out, fpn_out, 3, bias=True)
    for i in (1, 2):
        g.sd[f"FPN.smooth_conv.{i}.weight"] = g.sd["FPN.smooth_conv.0.weight"]
        g.sd[f"FPN.smooth_conv.{i}.bias"] = g.sd["FPN.smooth_conv.0.bias"]
    g.conv("FPN.conv_fusion.0", fpn_out, 4 * fpn_out, 3)
    g.bn("FPN.conv_fusion.1", fpn_out, beta0=0.0)
    g.conv("head", num_classes, fpn_out, 3, bias=True)
    return g.sd


def _xception_block(g, prefix, cin, cout, stride, exit_flow=False, use_1st_relu=True):
    """Block (deeplabv3_plus.py:89-132): registration order skip, skipbn, then rep.* with ReLU modules occupying indices."""
    if cin != cout or stride != 1:
        g.conv(prefix + "skip", cout, cin, 1)
        g.bn(prefix + "skipbn", cout)
    if exit_flow:
        units = [(cin, cin), (cin, cout), (cout, cout)]
    else:
        units = [(cin, cout), (cout, cout), (cout, cout)]
    idx = 0 if not use_1st_relu else 1
    for a, b in units:
        g.sep(f"{prefix}rep.{idx}", a, b)
        g.bn(f"{prefix}rep.{idx + 1}", b)
        idx += 3


def deeplab_xception_state_dict(num_classes, seed=0, randomize_bn=False, output_stride=16):
    """Keys/shapes of models.DeepLab(num_classes, backbone='xception', pretrained=False).state_dict()
    (deeplabv3_plus.py:134-199,253-330): 848 entries."""
    g = _Gen(seed, randomize_bn)
    g.conv("backbone.conv1", 32, 3, 3)
    g.bn("backbone.bn1", 32)
    g.conv("backbone.conv2", 64, 32, 3)
    g.bn("backbone.bn2", 64)
    _xception_block(g, "backbone.block1.", 64, 128, 2, use_1st_relu=False)
    _xception_block(g, "backbone.block2.", 128, 256, 2)
    _xception_block(g, "backbone.block3.", 256, 728, 2 if output_stride == 16 else 1)
    for i in range(16):
        _xception_block(g, f"backbone.block{i + 4}.", 728, 728, 1)
    _xception_block(g, "backbone.block20.", 728, 1024, 1, exit_flow=True)
    g.sep("backbone.conv3", 1024, 1536)
    g.bn("backbone.bn3", 1536)
    g.sep("backbone.conv4", 1536, 1536)
    g.bn("backbone.bn4", 1536)
    g.sep("backbone.conv5", 1536, 2048)
    g.bn("backbone.bn5", 2048)
    for i, k in zip((1, 2, 3, 4), (1, 3, 3, 3)):
        g.conv(f"ASSP.aspp{i}.0", 256, 2048, k)
        g.bn(f"ASSP.aspp{i}.1", 256)
    g.conv("ASSP.avg_pool.1", 256, 2048, 1)
    g.bn("ASSP.avg_pool.2", 256)
    g.conv("ASSP.conv1", 256, 1280, 1)
    g.bn("ASSP.bn1", 256)
    g.conv("decoder.conv1", 48, 128, 1)
    g.bn("decoder.bn1", 48)
    g.conv("decoder.output.0", 256, 304, 3)
    g.bn("decoder.output.1", 256)
    g.conv("decoder.output.3", 256, 256, 3)
    g.bn("decoder.output.4", 256)
    g.conv("decoder.output.7", num_classes, 256, 1, bias=True)
    return g.sd
