#!/usr/bin/env python
"""Where does the engine's eval-mode forward drift from the fp32 oracle?  DeepLabV3+/ResNet-101 with the initialisers'
running statistics (BatchNorm = identity, an un-normalised residual trunk whose activations grow with depth), block by
block (each side runs ITS OWN chain): norm ratio and relative L2 error of the engine and of the ATen-bf16 control against
the fp32 oracle after the stem and after every Bottleneck.

    python tools/eval_init_probe.py [--size 513] [--batch 2]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-segmentation_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
import seg_b200  # noqa: E402
from oracle import models as om, synth, weights  # noqa: E402
from seg_b200.engine import Act  # noqa: E402


def chain(osd, x):
    outs = []
    h = om._conv(osd, "backbone.layer0.0", x, 2, 3)
    outs.append(("stem.conv", h))
    h = F.relu(om._bn(osd, "backbone.layer0.1", h, False))
    outs.append(("stem.bn_relu", h))
    h = F.max_pool2d(h, 3, 2, 1)
    outs.append(("stem.maxpool", h))
    cfg = {1: (1, 1), 2: (2, 1), 3: (2, 1), 4: (1, 2)}
    for li in (1, 2, 3, 4):
        stride, dil = cfg[li]
        for b in range(weights.RESNET_LAYERS["resnet101"][li - 1]):
            h = om._bottleneck(osd, f"backbone.layer{li}.{b}.", h, stride if b == 0 else 1, dil, False)
            outs.append((f"layer{li}.{b}", h))
    return outs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=513)
    ap.add_argument("--batch", type=int, default=2)
    a = ap.parse_args()
    sd = weights.deeplab_resnet_state_dict(19, "resnet101", seed=0)
    x, _ = synth.make_batch(a.batch, a.size, a.size, 19, 255, seed=9001)
    with torch.no_grad():
        ref = chain(om.clone_sd(sd), x)
        dsd = {k: v.cuda() for k, v in sd.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ctl = chain(dsd, x.cuda())
        m = seg_b200.DeepLab(19, backbone="resnet101", pretrained=False)
        m.load_state_dict(sd)
        m = m.cuda().eval()
        tape = m._new_tape(False, False)
        bb = m.backbone
        eng = []
        y, _ = tape.conv(x.cuda().contiguous().float(), m._spec("backbone.layer0.0", bb.layer0[0]), want_stats=True)
        eng.append(("stem.conv", y.t))
        h = tape.bn_act(y, bb.layer0[1], None)
        eng.append(("stem.bn_relu", h.t))
        h = tape.maxpool(h)
        eng.append(("stem.maxpool", h.t))
        for li in (1, 2, 3, 4):
            for bi, blk in enumerate(getattr(bb, f"layer{li}")):
                h = m._block(tape, h, f"backbone.layer{li}.{bi}.", blk)
                eng.append((f"layer{li}.{bi}", h.t))
        torch.cuda.synchronize()
    print(f"{'after':14s} {'|ref|rms':>10s} | engine: norm ratio-1   rel L2 | control: norm ratio-1   rel L2")
    for (n, r), (_, c), (_, e) in zip(ref, ctl, eng):
        r = r.double()
        e = e.float().permute(0, 3, 1, 2).double().cpu()
        c = c.float().double().cpu()
        print(f"{n:14s} {r.pow(2).mean().sqrt().item():10.3e} | {e.norm().item() / r.norm().item() - 1:+.3e}  {((e - r).norm() / r.norm()).item():.3e}"
              f" | {c.norm().item() / r.norm().item() - 1:+.3e}  {((c - r).norm() / r.norm()).item():.3e}")


if __name__ == "__main__":
    main()
