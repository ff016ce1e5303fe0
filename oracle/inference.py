"""CPU restatement of the reference's test-time augmentation (inference.py:20-79): TEST INFRASTRUCTURE ONLY — imported
by tests/ (the checker), never by the product path.

The arithmetic of this path lives in un-vendored third-party code: scipy.ndimage.zoom (requirements.txt pins scipy),
torch.nn.Upsample / F.pad on the CPU and numpy float64 accumulation.  The restatement calls the same library entry
points the reference calls, in the same order, and is pinned by tests/golden/inference.npz, which
oracle/make_golden_inference.py produced by importing the reference's own functions in this container.
`model` is any callable NCHW float32 tensor -> NCHW float32 tensor on `device`.
"""
from math import ceil

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from scipy import ndimage


def pad_image(img, target_size):  # inference.py:20-24
    rows_to_pad = max(target_size[0] - img.shape[2], 0)
    cols_to_pad = max(target_size[1] - img.shape[3], 0)
    return F.pad(img, (0, cols_to_pad, 0, rows_to_pad), "constant", 0)


def sliding_predict(model, image, num_classes, flip=True):  # inference.py:26-56
    size = image.shape
    tile = (int(size[2] // 2.5), int(size[3] // 2.5))
    stride = ceil(tile[0] * (1 - 1 / 3))
    rows = int(ceil((size[2] - tile[0]) / stride) + 1)
    cols = int(ceil((size[3] - tile[1]) / stride) + 1)
    total = np.zeros((num_classes, size[2], size[3]))
    count = np.zeros((size[2], size[3]))
    for row in range(rows):
        for col in range(cols):
            x_min, y_min = int(col * stride), int(row * stride)
            x_max, y_max = min(x_min + tile[1], size[3]), min(y_min + tile[0], size[2])
            img = image[:, :, y_min:y_max, x_min:x_max]
            padded = pad_image(img, tile)
            pred = model(padded)
            if flip:
                pred = 0.5 * (model(padded.flip(-1)).flip(-1) + pred)
            pred = pred[:, :, :img.shape[2], :img.shape[3]]
            count[y_min:y_max, x_min:x_max] += 1
            total[:, y_min:y_max, x_min:x_max] += pred.data.cpu().numpy().squeeze(0)
    total /= count
    return total


def multi_scale_predict(model, image, scales, num_classes, device, flip=False):  # inference.py:58-79
    input_size = (image.size(2), image.size(3))
    upsample = nn.Upsample(size=input_size, mode="bilinear", align_corners=True)
    total = np.zeros((num_classes, image.size(2), image.size(3)))
    image = image.data.cpu().numpy()
    for scale in scales:
        scaled = ndimage.zoom(image, (1.0, 1.0, float(scale), float(scale)), order=1, prefilter=False)
        scaled = torch.from_numpy(scaled).to(device)
        pred = upsample(model(scaled).cpu())
        if flip:
            pred_f = upsample(model(scaled.flip(-1).to(device)).cpu())
            pred = 0.5 * (pred_f.flip(-1) + pred)
        total += pred.data.cpu().numpy().squeeze(0)
    total /= len(scales)
    return total


def toy_model(num_classes, seed=0):
    """Deterministic fp32 stand-in network for the golden vectors and the kernel-logic tests: 3x3 conv (3 -> C), tanh,
    1x1 conv — shape-generic, smooth, NOT left/right symmetric (so the flip branches matter).  Works on any device."""
    g = torch.Generator().manual_seed(seed)
    w1 = torch.randn(num_classes, 3, 3, 3, generator=g) * 0.4
    w2 = torch.randn(num_classes, num_classes, 1, 1, generator=g) * 0.7
    b = torch.randn(num_classes, generator=g) * 0.1

    def model(x):
        h = torch.tanh(F.conv2d(x, w1.to(x.device), padding=1))
        return F.conv2d(h, w2.to(x.device), b.to(x.device))

    return model
