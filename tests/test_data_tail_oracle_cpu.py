"""oracle/data.py (CPU restatement of base_dataset.py:93-123,129-136) against golden vectors produced by the reference's
own BaseDataSet.__getitem__ (oracle/make_golden_data.py), plus the host-side draw order of seg_b200.data."""
import os
import random

import numpy as np
import torch

from oracle import data as od

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data_tail.npz")


def test_oracle_tail_is_bit_exact_against_reference():
    g = np.load(GOLD)
    crop = int(g["crop"])
    for i in range(int(g["n"])):
        y0, x0, flip = (int(v) for v in g[f"{i}/draw"])
        x, y = od.sample_tail(g[f"{i}/image"], g[f"{i}/label"], crop, y0, x0, bool(flip), g["mean"].tolist(), g["std"].tolist())
        assert torch.equal(x, torch.from_numpy(g[f"{i}/x"])), i
        assert torch.equal(y, torch.from_numpy(g[f"{i}/y"])), i


def test_draw_order_matches_reference():
    from seg_b200.data import draw_crop_flip
    g = np.load(GOLD)
    crop = int(g["crop"])
    for i in range(int(g["n"])):
        h, w = g[f"{i}/image"].shape[:2]
        random.seed(100 + i)  # the seed the golden generator gave the reference's __getitem__
        y0, x0, flip = draw_crop_flip(h, w, crop, flip=True)
        assert [y0, x0, int(flip)] == [int(v) for v in g[f"{i}/draw"]], i
