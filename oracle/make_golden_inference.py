#!/usr/bin/env python
"""Golden vectors for inference.py::multi_scale_predict / sliding_predict, produced by importing the REFERENCE
(read-only, /root/reference) in this container with a deterministic stand-in network (oracle.inference.toy_model).
Writes tests/golden/inference.npz.  Run:  python oracle/make_golden_inference.py"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SEG_REFERENCE_ROOT", "/root/reference")


def main():
    sys.path.insert(0, REF)
    for name in ("skimage", "skimage.filters"):  # utils/__init__ import chain needs it; unused on this path
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.gaussian = lambda *a, **k: None
            sys.modules[name] = m
    import inference as ref  # the reference's inference.py
    sys.path.insert(0, ROOT)
    from oracle.inference import toy_model
    torch.set_num_threads(4)
    C = 5
    model = toy_model(C, seed=3)
    g = torch.Generator().manual_seed(9300)
    rec = {"num_classes": np.asarray(C)}
    for tag, (H, W) in (("a", (37, 53)), ("b", (64, 48))):
        img = torch.randn(1, 3, H, W, generator=g)
        rec[f"{tag}/image"] = img.numpy()
        scales = [0.75, 1.0, 1.25, 1.5, 1.75, 2.0]
        rec[f"{tag}/scales"] = np.asarray(scales)
        with torch.no_grad():
            rec[f"{tag}/ms"] = ref.multi_scale_predict(model, img, scales, C, torch.device("cpu"), flip=False)
            rec[f"{tag}/ms_flip"] = ref.multi_scale_predict(model, img, scales, C, torch.device("cpu"), flip=True)
            rec[f"{tag}/slide"] = ref.sliding_predict(model, img, C, flip=False)
            rec[f"{tag}/slide_flip"] = ref.sliding_predict(model, img, C, flip=True)
        for k in ("ms", "ms_flip", "slide", "slide_flip"):
            print(tag, k, rec[f"{tag}/{k}"].shape, float(np.abs(rec[f"{tag}/{k}"]).mean()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "inference.npz"), **rec)


if __name__ == "__main__":
    main()
