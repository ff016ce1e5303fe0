#!/usr/bin/env python
"""Golden vectors for utils/metrics.py::eval_metrics, produced by importing the REFERENCE (read-only, /root/reference) in
this container.  Writes tests/golden/metrics.npz.  Run:  python oracle/make_golden_metrics.py"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SEG_REFERENCE_ROOT", "/root/reference")


def main():
    sys.path.insert(0, REF)
    for name in ("skimage", "skimage.filters"):  # utils/__init__ import chain needs it; not used by metrics
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.gaussian = lambda *a, **k: None
            sys.modules[name] = m
    from utils.metrics import eval_metrics
    sys.path.insert(0, ROOT)
    g = torch.Generator().manual_seed(9005)
    rec = {}
    for tag, C, ignore, shape in (("c7", 7, 255, (2, 12, 13)), ("c19", 19, 255, (3, 33, 35)), ("c150", 150, -1, (2, 24, 20))):
        logits = torch.randn(shape[0], C, shape[1], shape[2], generator=g) * 2
        target = torch.randint(0, C, shape, generator=g)
        target[:, :2, :] = ignore
        target[:, :, -1] = ignore
        # make a good part of the predictions correct so that every counter is exercised
        hot = torch.rand(shape, generator=g) < 0.6
        idx = target.clamp(0, C - 1)
        logits.scatter_add_(1, idx.unsqueeze(1), (hot.float() * 6).unsqueeze(1))
        out = eval_metrics(logits, target, C)
        rec[f"{tag}/logits"] = logits.numpy()
        rec[f"{tag}/target"] = target.numpy()
        rec[f"{tag}/correct"] = np.asarray(out[0], dtype=np.int64)
        rec[f"{tag}/labeled"] = np.asarray(out[1], dtype=np.int64)
        rec[f"{tag}/inter"] = np.asarray(out[2], dtype=np.float32)
        rec[f"{tag}/union"] = np.asarray(out[3], dtype=np.float32)
        print(tag, int(out[0]), int(out[1]), float(out[2].sum()), float(out[3].sum()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "metrics.npz"), **rec)


if __name__ == "__main__":
    main()
