#!/usr/bin/env python
"""Golden vectors for the sample-pipeline tail, produced by the REFERENCE's own BaseDataSet.__getitem__
(/root/reference/base/base_dataset.py, read-only) on synthetic uint8 samples with scale / rotate / blur off, so the only
operations are the ones seg_augment_batch_u8 replaces: pad, crop, flip, ToTensor, Normalize.  The crop / flip draws are
recovered by replaying Python's `random` from the same seed.  Writes tests/golden/data_tail.npz.
Run:  python oracle/make_golden_data.py"""
import os
import random
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SEG_REFERENCE_ROOT", "/root/reference")
MEAN, STD = [0.45734706, 0.43338275, 0.40058118], [0.23965294, 0.23532275, 0.2398498]  # dataloaders/voc.py
CROP = 48


def main():
    sys.path.insert(0, REF)
    for name in ("skimage", "skimage.filters"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.gaussian = lambda *a, **k: None
            sys.modules[name] = m
    from base.base_dataset import BaseDataSet

    rs = np.random.RandomState(9500)
    sizes = [(33, 70), (48, 48), (90, 41), (64, 80), (20, 30), (75, 49)]
    samples = [(rs.randint(0, 256, (h, w, 3)).astype(np.uint8), rs.randint(0, 21, (h, w)).astype(np.int32)) for h, w in sizes]

    class Synth(BaseDataSet):
        def _set_files(self):
            self.files = list(range(len(samples)))

        def _load_data(self, index):
            im, lb = samples[index]
            return im.astype(np.float32), lb, str(index)  # loaders hand float32 images to _augmentation (dataloaders/voc.py:44)

    ds = Synth(root=None, split="train", mean=MEAN, std=STD, base_size=None, augment=True, val=False, crop_size=CROP,
               scale=False, flip=True, rotate=False, blur=False)
    rec = {"mean": np.asarray(MEAN), "std": np.asarray(STD), "crop": np.asarray(CROP), "n": np.asarray(len(samples))}
    for i, (im, lb) in enumerate(samples):
        random.seed(100 + i)
        x, y = ds[i]
        random.seed(100 + i)  # replay the draws of base_dataset.py:107-121
        h, w = im.shape[:2]
        ph, pw = max(h, CROP), max(w, CROP)
        y0 = random.randint(0, ph - CROP)
        x0 = random.randint(0, pw - CROP)
        flip = random.random() > 0.5
        rec[f"{i}/image"], rec[f"{i}/label"] = im, lb
        rec[f"{i}/draw"] = np.asarray([y0, x0, int(flip)])
        rec[f"{i}/x"], rec[f"{i}/y"] = x.numpy(), y.numpy()
        print(i, im.shape, (y0, x0, flip), x.shape, y.shape, float(x.mean()))
    # ---- with the random-scale resize in front (base_dataset.py:66-75): base_size set, scale=True
    BASE = 60
    ds2 = Synth(root=None, split="train", mean=MEAN, std=STD, base_size=BASE, augment=True, val=False, crop_size=CROP,
                scale=True, flip=True, rotate=False, blur=False)
    rec["base_size"] = np.asarray(BASE)
    for i, (im, lb) in enumerate(samples):
        random.seed(200 + i)
        x, y = ds2[i]
        random.seed(200 + i)  # replay: long side, crop row, crop column, flip
        h0, w0 = im.shape[:2]
        longside = random.randint(int(BASE * 0.5), int(BASE * 2.0))
        h, w = (longside, int(1.0 * longside * w0 / h0 + 0.5)) if h0 > w0 else (int(1.0 * longside * h0 / w0 + 0.5), longside)
        ph, pw = max(h, CROP), max(w, CROP)
        y0 = random.randint(0, ph - CROP)
        x0 = random.randint(0, pw - CROP)
        flip = random.random() > 0.5
        rec[f"s{i}/draw"] = np.asarray([h, w, y0, x0, int(flip)])
        rec[f"s{i}/x"], rec[f"s{i}/y"] = x.numpy(), y.numpy()
        print("scaled", i, im.shape, (h, w), (y0, x0, flip), float(x.mean()))
    # ---- scale + rotate (base_dataset.py:77-83), the shipped config.json's setting
    ds3 = Synth(root=None, split="train", mean=MEAN, std=STD, base_size=BASE, augment=True, val=False, crop_size=CROP,
                scale=True, flip=True, rotate=True, blur=False)
    for i, (im, lb) in enumerate(samples):
        random.seed(300 + i)
        x, y = ds3[i]
        random.seed(300 + i)  # replay: long side, angle, crop row, crop column, flip
        h0, w0 = im.shape[:2]
        longside = random.randint(int(BASE * 0.5), int(BASE * 2.0))
        h, w = (longside, int(1.0 * longside * w0 / h0 + 0.5)) if h0 > w0 else (int(1.0 * longside * h0 / w0 + 0.5), longside)
        angle = random.randint(-10, 10)
        ph, pw = max(h, CROP), max(w, CROP)
        y0 = random.randint(0, ph - CROP)
        x0 = random.randint(0, pw - CROP)
        flip = random.random() > 0.5
        rec[f"r{i}/draw"] = np.asarray([h, w, angle, y0, x0, int(flip)])
        rec[f"r{i}/x"], rec[f"r{i}/y"] = x.numpy(), y.numpy()
        print("rotated", i, im.shape, (h, w), angle, (y0, x0, flip), float(x.mean()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "data_tail.npz"), **rec)


if __name__ == "__main__":
    main()
