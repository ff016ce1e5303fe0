#!/usr/bin/env python
"""Op-by-op, teacher-forced look inside ONE Bottleneck in eval mode (identity BatchNorm): each engine op is fed the engine's
own previous output and compared with the fp32 ATen result ON THAT SAME INPUT — norm ratio (a systematic shrink shows as a
negative number far above the fp32 noise), relative L2 error and the mean signed error.

    python tools/eval_block_probe.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-segmentation_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
import seg_b200  # noqa: E402
from oracle import models as om, synth, weights  # noqa: E402
from seg_b200 import ops  # noqa: E402
from seg_b200.engine import Act  # noqa: E402


def nchw(t):
    return t.float().permute(0, 3, 1, 2).contiguous()


def report(name, got_nhwc, ref_nchw):
    g = nchw(got_nhwc).double().cpu()
    r = ref_nchw.double().cpu()
    print(f"{name:34s} norm ratio-1 {g.norm().item() / r.norm().item() - 1:+.3e}   rel L2 {((g - r).norm() / r.norm()).item():.3e}   "
          f"mean signed (g-r)*sign(r)/mean|r| {(((g - r) * r.sign()).mean() / r.abs().mean()).item():+.3e}")


def main():
    sd = weights.deeplab_resnet_state_dict(19, "resnet101", seed=0)
    x, _ = synth.make_batch(2, 513, 513, 19, 255, seed=9001)
    m = seg_b200.DeepLab(19, backbone="resnet101", pretrained=False)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    dsd = {k: v.cuda() for k, v in sd.items()}
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        tape = m._new_tape(False, False)
        bb = m.backbone
        h = m._cbr(tape, x.cuda().contiguous().float(), "backbone.layer0.0", bb.layer0[0], bb.layer0[1])
        h = tape.maxpool(h)
        for name, blk, stride, dil in (("backbone.layer1.0.", bb.layer1[0], 1, 1), ("backbone.layer1.1.", bb.layer1[1], 1, 1)):
            print(f"--- {name} (input {tuple(h.t.shape)}) ---")
            xin = nchw(h.t)  # the engine's own input, exactly
            y1, _ = tape.conv(h, m._spec(name + "conv1", blk.conv1), want_stats=True)
            report("conv1 (1x1)", y1.t, F.conv2d(xin, dsd[name + "conv1.weight"].bfloat16().float()))
            a1 = tape.bn_act(y1, blk.bn1, None)
            report("bn1+relu (eval)", a1.t, F.relu(om._bn(dsd, name + "bn1", nchw(y1.t), False)))
            y2, _ = tape.conv(a1, m._spec(name + "conv2", blk.conv2), want_stats=True)
            report("conv2 (3x3)", y2.t, F.conv2d(nchw(a1.t), dsd[name + "conv2.weight"].bfloat16().float(), None, stride, dil, dil))
            a2 = tape.bn_act(y2, blk.bn2, None)
            report("bn2+relu (eval)", a2.t, F.relu(om._bn(dsd, name + "bn2", nchw(y2.t), False)))
            y3, _ = tape.conv(a2, m._spec(name + "conv3", blk.conv3), want_stats=True)
            report("conv3 (1x1)", y3.t, F.conv2d(nchw(a2.t), dsd[name + "conv3.weight"].bfloat16().float()))
            r = h
            rref = xin
            if blk.downsample is not None:
                yd, _ = tape.conv(h, m._spec(name + "downsample.0", blk.downsample[0]), want_stats=True)
                report("downsample conv (1x1)", yd.t, F.conv2d(xin, dsd[name + "downsample.0.weight"].bfloat16().float(), None, stride))
                r = tape.bn_act(yd, blk.downsample[1], None, relu=False)
                report("downsample bn (eval)", r.t, om._bn(dsd, name + "downsample.1", nchw(yd.t), False))
                rref = nchw(r.t)
            a3 = tape.bn_act(y3, blk.bn3, None, relu=True, res=r)
            report("bn3 + residual + relu (eval)", a3.t, F.relu(om._bn(dsd, name + "bn3", nchw(y3.t), False) + rref))
            h = a3
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()
