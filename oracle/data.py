"""CPU restatement of the tail of the reference's sample pipeline — base/base_dataset.py:93-123 (zero-pad to the crop
size, crop, horizontal flip) and :129-136 (label -> int64, ToTensor, Normalize).  TEST INFRASTRUCTURE ONLY.
Pinned by tests/golden/data_tail.npz (oracle/make_golden_data.py runs the reference's own BaseDataSet.__getitem__)."""
import numpy as np
import torch


def sample_tail(image, label, crop_size, y0, x0, flip, mean, std):
    """image uint8 [h,w,3], label [h,w] -> (fp32 [3,crop,crop], int64 [crop,crop])"""
    h, w, _ = image.shape
    pad_h, pad_w = max(crop_size - h, 0), max(crop_size - w, 0)
    if pad_h > 0 or pad_w > 0:  # cv2.copyMakeBorder(top=0, bottom=pad_h, left=0, right=pad_w, BORDER_CONSTANT, value=0)
        image = np.pad(image, ((0, pad_h), (0, pad_w), (0, 0)), mode="constant", constant_values=0)
        label = np.pad(label, ((0, pad_h), (0, pad_w)), mode="constant", constant_values=0)
    image = image[y0:y0 + crop_size, x0:x0 + crop_size]
    label = label[y0:y0 + crop_size, x0:x0 + crop_size]
    if flip:
        image = np.fliplr(image).copy()
        label = np.fliplr(label).copy()
    lab = torch.from_numpy(np.array(label, dtype=np.int32)).long()
    t = torch.from_numpy(np.ascontiguousarray(np.uint8(image))).permute(2, 0, 1).contiguous().to(torch.float32).div(255)  # ToTensor
    m = torch.as_tensor(mean, dtype=torch.float32).view(-1, 1, 1)
    s = torch.as_tensor(std, dtype=torch.float32).view(-1, 1, 1)
    return t.sub_(m).div_(s), lab  # Normalize


def _cv_coeffs(dst, src, clamp):
    """cv::resize INTER_LINEAR source index / weight per destination index (modules/imgproc/src/resize.cpp): float64
    coordinate rounded to float32; along x the weight is zeroed where the index is clamped, along y only the ROW INDEX is
    clipped (the weight is kept)."""
    scale = 1.0 / (dst / src)
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f)
    f = (f - s).astype(np.float32)
    s = s.astype(np.int64)
    if clamp:
        lo, hi = s < 0, s >= src - 1
        f[lo | hi] = 0
        s[lo] = 0
        s[hi] = src - 1
    return s, (np.float32(1.0) - f).astype(np.float32), f


def cv_resize_linear_f32(img, w, h):
    """cv2.resize(img.astype(float32), (w, h), INTER_LINEAR) as OpenCV's own code computes it (fp32 products and sums in
    this order, no fma).  Bit-exact against cv2 with IPP disabled (cv2.ipp.setUseIPP(False)); the IPP float kernels that
    opencv-python wheels dispatch to by default differ by <= 3e-3 (tests/test_data_tail_oracle_cpu.py measures both)."""
    S = img.astype(np.float32)
    sh, sw = S.shape[:2]
    xi, ax0, ax1 = _cv_coeffs(w, sw, True)
    yi, ay0, ay1 = _cv_coeffs(h, sh, False)
    x1 = np.minimum(xi + 1, sw - 1)
    H = S[:, xi] * ax0[None, :, None] + S[:, x1] * ax1[None, :, None]
    y0, y1 = np.clip(yi, 0, sh - 1), np.clip(yi + 1, 0, sh - 1)
    return (H[y0] * ay0[:, None, None] + H[y1] * ay1[:, None, None]).astype(np.float32)


def cv_resize_nearest(lbl, w, h):
    """cv2.resize(lbl, (w, h), INTER_NEAREST): index = min(floor(d * scale), src - 1), scale in float64."""
    sh, sw = lbl.shape[:2]
    xs = np.minimum(np.floor(np.arange(w) * (1.0 / (w / sw))).astype(np.int64), sw - 1)
    ys = np.minimum(np.floor(np.arange(h) * (1.0 / (h / sh))).astype(np.int64), sh - 1)
    return lbl[ys][:, xs]


def sample_scale_tail(image, label, h, w, crop_size, y0, x0, flip, mean, std):
    """base_dataset.py:66-75 (random-scale resize to h x w) followed by the tail above; the float image is truncated to
    uint8 where the reference does it (np.uint8(image), base_dataset.py:133) — truncation commutes with pad/crop/flip."""
    img = np.uint8(cv_resize_linear_f32(image, w, h))
    lab = cv_resize_nearest(np.asarray(label), w, h)
    return sample_tail(img, lab, crop_size, y0, x0, flip, mean, std)
