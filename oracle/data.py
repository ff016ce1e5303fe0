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


def cv_rotation_matrix(center, angle_deg, scale=1.0):
    """cv2.getRotationMatrix2D (modules/imgproc/src/imgwarp.cpp): float64 throughout."""
    import math
    a = angle_deg * (math.pi / 180.0)  # angle *= CV_PI/180
    alpha, beta = math.cos(a) * scale, math.sin(a) * scale
    cx, cy = float(np.float32(center[0])), float(np.float32(center[1]))  # Point2f
    return np.array([[alpha, beta, (1 - alpha) * cx - beta * cy], [-beta, alpha, beta * cx + (1 - alpha) * cy]], dtype=np.float64)


def cv_warp_affine(img, M, w, h, linear):
    """cv2.warpAffine(img, M, (w, h), flags=INTER_LINEAR | INTER_NEAREST) with the default constant-0 border, as
    base_dataset.py:77-83 calls it (float32 image / int32 label).  OpenCV inverts M in float64, walks the destination with
    FIXED-POINT source coordinates (10 fractional bits: per-column terms round(M00*x*1024), per-row terms
    round((M01*y + M02)*1024) + rounding offset), and for INTER_LINEAR keeps 5 fractional bits (a 1/32-pixel grid) whose
    bilinear weights are float32 products (1-fy)(1-fx), (1-fy)fx, fy(1-fx), fy*fx; taps outside the image contribute 0.
    Bit-exact against cv2 with IPP on or off (tests/test_data_tail_oracle_cpu.py)."""
    AB_BITS, INTER_BITS = 10, 5
    AB_SCALE, TAB = 1 << AB_BITS, 1 << INTER_BITS
    M = np.asarray(M, dtype=np.float64)
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22, A12, A21 = M[1, 1] * D, M[0, 0] * D, -M[0, 1] * D, -M[1, 0] * D
    b1, b2 = -A11 * M[0, 2] - A12 * M[1, 2], -A21 * M[0, 2] - A22 * M[1, 2]
    sh, sw = img.shape[:2]
    xs, ys = np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64)
    rnd = lambda v: np.rint(v).astype(np.int64)  # noqa: E731  saturate_cast<int>(double): round half to even
    delta = AB_SCALE // TAB // 2 if linear else AB_SCALE // 2
    X = (rnd((A12 * ys + b1) * AB_SCALE) + delta)[:, None] + rnd(A11 * xs * AB_SCALE)[None, :]
    Y = (rnd((A22 * ys + b2) * AB_SCALE) + delta)[:, None] + rnd(A21 * xs * AB_SCALE)[None, :]

    def tap(yy, xx):
        ok = (xx >= 0) & (xx < sw) & (yy >= 0) & (yy < sh)
        v = img[np.clip(yy, 0, sh - 1), np.clip(xx, 0, sw - 1)]
        return np.where(ok[..., None] if img.ndim == 3 else ok, v, 0)

    if not linear:
        return tap(Y >> AB_BITS, X >> AB_BITS).astype(img.dtype)
    Xf, Yf = X >> (AB_BITS - INTER_BITS), Y >> (AB_BITS - INTER_BITS)
    xi, yi = Xf >> INTER_BITS, Yf >> INTER_BITS
    fx = (Xf & (TAB - 1)).astype(np.float32) / np.float32(TAB)
    fy = (Yf & (TAB - 1)).astype(np.float32) / np.float32(TAB)
    ex = (lambda a: a[..., None]) if img.ndim == 3 else (lambda a: a)
    one = np.float32(1)
    w00, w01, w10, w11 = ex((one - fy) * (one - fx)), ex((one - fy) * fx), ex(fy * (one - fx)), ex(fy * fx)
    f = img.astype(np.float32)
    sav, img = img, f
    out = tap(yi, xi) * w00 + tap(yi, xi + 1) * w01 + tap(yi + 1, xi) * w10 + tap(yi + 1, xi + 1) * w11
    img = sav
    return out.astype(np.float32)


def sample_scale_tail(image, label, h, w, crop_size, y0, x0, flip, mean, std, angle=None):
    """base_dataset.py:66-75 (random-scale resize to h x w), optionally :77-83 (rotation by `angle` degrees about the centre
    of the resized image), then the tail above; the float image is truncated to uint8 where the reference does it
    (np.uint8(image), base_dataset.py:133) — truncation commutes with pad/crop/flip."""
    img = cv_resize_linear_f32(image, w, h)
    lab = cv_resize_nearest(np.asarray(label), w, h)
    if angle is not None:
        M = cv_rotation_matrix((w / 2, h / 2), angle, 1.0)
        img = cv_warp_affine(img, M, w, h, linear=True)
        lab = cv_warp_affine(lab, M, w, h, linear=False)
    return sample_tail(np.uint8(img), lab, crop_size, y0, x0, flip, mean, std)
