#!/usr/bin/env python
"""Debug tool: find kernels whose result depends on uninitialised memory.

Every `torch.empty` the engine issues is replaced by a NaN-filled tensor (integer tensors: 0x7f bytes), every wrapper in
seg_b200.ops is followed by a NaN check of the tensors it returned / was handed as `out=`; the first ops that produce NaN
in the LOGICAL region of a tensor (padding outside a view does not count) are listed.  A clean engine prints none and
the poisoned run's logits / gradients equal the unpoisoned run's.

    python tools/poison_probe.py [deeplab|pspnet|upernet|xception] [size] [backbone]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-segmentation_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

import seg_b200  # noqa: E402
from seg_b200 import engine, nets, ops  # noqa: E402
from seg_b200 import losses as plosses  # noqa: E402
from oracle import synth, weights  # noqa: E402

_real_empty = torch.empty
POISON = {"on": False}
REPORT = []


def poisoned_empty(*a, **k):
    t = _real_empty(*a, **k)
    if POISON["on"] and t.is_cuda and t.numel():
        if t.dtype.is_floating_point:
            t.fill_(float("nan"))
        else:
            t.view(torch.uint8).fill_(0x7F)
    return t


class _TorchProxy:
    """`torch` as seen by the engine modules: everything forwards to torch, `empty` poisons."""

    def __getattr__(self, name):
        return getattr(torch, name)

    empty = staticmethod(poisoned_empty)


def _tensors(obj):
    if isinstance(obj, torch.Tensor):
        yield obj
    elif isinstance(obj, (tuple, list)):
        for o in obj:
            yield from _tensors(o)


def wrap(name, fn):
    def inner(*a, **k):
        r = fn(*a, **k)
        if POISON["on"]:
            seen = []
            for label, t in [("ret", t) for t in _tensors(r)] + [(f"kw:{kk}", v) for kk, v in k.items() if isinstance(v, torch.Tensor)]:
                if t.is_cuda and t.dtype.is_floating_point and t.numel():
                    n = int(torch.isnan(t).sum())
                    if n:
                        seen.append(f"{label}{tuple(t.shape)}:{n}/{t.numel()}")
            if seen and len(REPORT) < 60:
                shapes = [tuple(t.shape) for t in _tensors(a)]
                REPORT.append(f"{name} args{shapes} -> NaN in {seen}")
        return r
    return inner


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "deeplab"
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 65
    backbone = sys.argv[3] if len(sys.argv) > 3 else "resnet14"
    nc = 7
    if kind == "deeplab":
        sd = weights.deeplab_resnet_state_dict(nc, backbone, seed=21, randomize_bn=True)
        m = seg_b200.DeepLab(nc, backbone=backbone, output_stride=16)
    elif kind == "xception":
        sd = weights.deeplab_xception_state_dict(nc, seed=21, randomize_bn=True, output_stride=16)
        m = seg_b200.DeepLab(nc, backbone="xception", output_stride=16)
    elif kind == "upernet":
        sd = weights.upernet_state_dict(nc, backbone, seed=21, randomize_bn=True)
        m = seg_b200.UperNet(nc, backbone=backbone)
    else:
        sd = weights.pspnet_state_dict(nc, backbone, seed=21, randomize_bn=True)
        m = seg_b200.PSPNet(nc, backbone=backbone)
    m.load_state_dict(sd, strict=True)
    m.engine_dropout = False
    m = m.cuda().train()
    crit = seg_b200.CrossEntropyLoss2d(ignore_index=255)
    x, y = synth.make_batch(2, size, size, nc, 255, seed=9100)
    xd, yd = x.cuda(), y.cuda()
    snap = {k: v.detach().clone() for k, v in m.state_dict().items()}

    def run():
        with torch.no_grad():
            for k, v in m.state_dict().items():
                v.copy_(snap[k])
        m.zero_grad(set_to_none=True)
        out = m(xd)
        outs = out if isinstance(out, tuple) else (out,)
        loss = sum(crit(o, yd) for o in outs)
        loss.backward()
        torch.cuda.synchronize()
        g = torch.cat([p.grad.reshape(-1) for p in m.parameters() if p.grad is not None])
        return outs[0].detach().clone(), float(loss), g.clone()

    o0, l0, g0 = run()
    o1, l1, g1 = run()
    for name in dir(ops):
        fn = getattr(ops, name)
        if callable(fn) and getattr(fn, "__module__", None) == ops.__name__ and not name.startswith("_"):
            setattr(ops, name, wrap(name, fn))
    proxy = _TorchProxy()
    for mod in (ops, engine, nets, plosses):
        mod.torch = proxy
    POISON["on"] = True
    o2, l2, g2 = run()
    POISON["on"] = False

    def rel(a, b):
        return ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-12)).item()

    print(f"[{kind}/{backbone}@{size}] clean-vs-clean: logits {rel(o1, o0):.3e} loss {l0:.6f}/{l1:.6f} grads {rel(g1, g0):.3e}")
    print(f"[{kind}/{backbone}@{size}] poisoned-vs-clean: logits {rel(o2, o0):.3e} loss {l2:.6f} grads {rel(g2, g0):.3e} "
          f"NaN logits {int(torch.isnan(o2).sum())} NaN grads {int(torch.isnan(g2).sum())}")
    for line in REPORT:
        print("  ", line)
    if not REPORT:
        print("   no op produced NaN in a logical region")


if __name__ == "__main__":
    main()
