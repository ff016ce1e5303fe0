"""oracle/inference.py (CPU restatement of inference.py:20-79) against the golden vectors the reference's own functions
produced (oracle/make_golden_inference.py)."""
import os

import numpy as np
import torch

from oracle import inference as oi

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "inference.npz")


def test_oracle_reproduces_reference_inference():
    g = np.load(GOLD)
    C = int(g["num_classes"])
    model = oi.toy_model(C, seed=3)
    for tag in ("a", "b"):
        img = torch.from_numpy(g[f"{tag}/image"])
        scales = [float(s) for s in g[f"{tag}/scales"]]
        with torch.no_grad():
            got = {"ms": oi.multi_scale_predict(model, img, scales, C, torch.device("cpu"), flip=False),
                   "ms_flip": oi.multi_scale_predict(model, img, scales, C, torch.device("cpu"), flip=True),
                   "slide": oi.sliding_predict(model, img, C, flip=False),
                   "slide_flip": oi.sliding_predict(model, img, C, flip=True)}
        for k, v in got.items():
            ref = g[f"{tag}/{k}"]
            assert v.shape == ref.shape
            # same library calls in the same order: bit-identical on the generating host; 1e-6 leaves room for another
            # host's conv kernels in the stand-in network
            assert np.abs(v - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max()), (tag, k, np.abs(v - ref).max())


def test_sliding_window_geometry_matches_reference_formula():
    from seg_b200.inference import sliding_windows
    for (H, W) in ((37, 53), (64, 48), (513, 513), (1024, 2048)):
        tile, wins = sliding_windows(H, W)
        assert tile == (int(H // 2.5), int(W // 2.5))
        cover = np.zeros((H, W), dtype=np.int32)
        for (y0, y1, x0, x1) in wins:
            assert 0 <= y0 < y1 <= H and 0 <= x0 < x1 <= W and y1 - y0 <= tile[0] and x1 - x0 <= tile[1]
            cover[y0:y1, x0:x1] += 1
        assert cover.min() >= 1, "every pixel is predicted at least once (the reference divides by this count)"
