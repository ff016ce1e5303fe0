"""Device implementation of the reference's test-time augmentation (seg_b200.inference: multi-scale + flip, sliding
window, label map) against the CPU oracle (oracle/inference.py, pinned to the reference by tests/golden/inference.npz)
and against the golden vectors themselves.  The stand-in network runs on the GPU on both sides, so the comparison
isolates what seg_data.cu computes: the image pyramid, flips, up-sampling, accumulation and the arg-max."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import inference as oi
from oracle import synth, weights

if torch.cuda.is_available():
    import seg_b200
    from seg_b200 import inference as di
    from seg_b200 import ops

if torch.cuda.is_available():  # the fp32 stand-in network must not run its convolutions in TF32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "inference.npz")
TOL = 2e-5  # fp32 accumulation over scales / windows against the reference's float64 numpy accumulation


def rel(a, b):
    return float(np.abs(a - b).max() / max(1e-12, np.abs(b).max()))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_device_tta_matches_reference_golden(tag, gpu_out_dir):
    g = np.load(GOLD)
    C = int(g["num_classes"])
    model = oi.toy_model(C, seed=3)
    img = torch.from_numpy(g[f"{tag}/image"]).cuda()
    scales = [float(s) for s in g[f"{tag}/scales"]]
    got = {"ms": di.multi_scale_predict(model, img, scales, C, flip=False),
           "ms_flip": di.multi_scale_predict(model, img, scales, C, flip=True),
           "slide": di.sliding_predict(model, img, C, flip=False),
           "slide_flip": di.sliding_predict(model, img, C, flip=True)}
    for k, v in got.items():
        ref = g[f"{tag}/{k}"]
        e = rel(v.cpu().double().numpy(), ref)
        lab = di.predict_labels(v).cpu().numpy()
        ref_lab = torch.softmax(torch.from_numpy(ref), dim=0).argmax(0).numpy()  # inference.py:156
        agree = float((lab == ref_lab).mean())
        msg = f"[inference {tag}/{k}] scores relerr {e:.2e}  label agreement {agree:.5f}"
        print(msg)
        with open(os.path.join(gpu_out_dir, "model_parity.txt"), "a") as f:
            f.write(msg + "\n")
        assert e < TOL, msg
        assert agree > 0.995, msg  # near-ties between two classes may flip within TOL


def test_resize_flip_window_kernels_vs_aten():
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(2, 3, 29, 41, generator=gen).cuda()
    for (Hd, Wd) in ((29, 41), (44, 30), (13, 97)):
        for ac in (True, False):
            ref = torch.nn.functional.interpolate(x, size=(Hd, Wd), mode="bilinear", align_corners=ac)
            got = ops.resize_nchw(x, Hd, Wd, align_corners=ac)
            assert (got - ref).abs().max().item() < 2e-6
            got_f = ops.resize_nchw(x, Hd, Wd, align_corners=ac, flip_x=True)
            assert (got_f - ref.flip(-1)).abs().max().item() < 2e-6
    assert torch.equal(ops.resize_nchw(x, 29, 41, flip_x=True), x.flip(-1)), "same-size resize must be an exact flip"
    acc = torch.ones(2, 3, 44, 30, device="cuda")
    ops.resize_nchw(x, 44, 30, alpha=0.25, out=acc, beta=2.0)
    ref = 2.0 + 0.25 * torch.nn.functional.interpolate(x, size=(44, 30), mode="bilinear", align_corners=True)
    assert (acc - ref).abs().max().item() < 2e-6
    dst = torch.zeros(2, 3, 50, 60, device="cuda")
    ops.window_add_nchw(x, dst, 7, 11, 20, 33, alpha=0.5)
    ops.window_add_nchw(x, dst, 7, 11, 20, 33, flip_x=True, alpha=0.5)
    ref = torch.zeros_like(dst)
    ref[:, :, 7:27, 11:44] = 0.5 * x[:, :, :20, :33] + 0.5 * x.flip(-1)[:, :, :20, :33]
    assert (dst - ref).abs().max().item() < 1e-6
    cnt = torch.randint(1, 4, (50, 60), generator=gen).float().cuda()
    assert torch.equal(ops.div_by_count_nchw(dst.clone(), cnt), dst / cnt)
    s = torch.randn(2, 7, 19, 23, generator=gen).cuda()
    s[:, 3] = s[:, 1]  # ties: the first maximum wins
    assert torch.equal(ops.argmax_nchw(s), s.argmax(1))


def test_zoom_mode_matches_scipy_including_its_black_edge():
    """seg_resize_nchw_f32 mode 2 against scipy.ndimage.zoom(order=1, prefilter=False) itself — including the size pairs
    (e.g. 48 -> 84) where scipy's float64 coordinate of the last column rounds past the last sample and the column is
    filled with 0 (mode='constant')."""
    from scipy import ndimage
    gen = torch.Generator().manual_seed(11)
    saw_black = False
    for (H, W) in ((64, 48), (37, 53), (97, 129)):
        x = torch.randn(1, 3, H, W, generator=gen) + 4.0
        for s in (0.75, 1.25, 1.5, 1.75, 2.0, 2.25):
            ref = ndimage.zoom(x.numpy(), (1.0, 1.0, float(s), float(s)), order=1, prefilter=False)
            got = ops.resize_nchw(x.cuda(), ref.shape[2], ref.shape[3], zoom=True).cpu().numpy()
            assert got.shape == ref.shape == (1, 3, int(round(H * s)), int(round(W * s)))
            assert np.abs(got - ref).max() < 2e-6, (H, W, s, np.abs(got - ref).max())
            saw_black |= bool((ref[..., :, -1] == 0).all() or (ref[..., -1, :] == 0).all())
    assert saw_black, "the parameter list is meant to include a pair that triggers scipy's constant-fill edge"


def test_engine_model_multi_scale_shapes_and_agreement(gpu_out_dir):
    """The real engine model under test-time augmentation: variable input sizes (scales 0.75 .. 1.5 of 97 x 129) run through
    the same kernels, and the device TTA agrees with the oracle TTA wrapped around the SAME model (bf16 noise only)."""
    nc = 7
    sd = weights.deeplab_resnet_state_dict(nc, "resnet14", seed=4, randomize_bn=True)
    m = seg_b200.DeepLab(nc, backbone="resnet14", output_stride=16)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    x, _ = synth.make_batch(1, 97, 129, nc, 255, seed=9400)
    scales = [0.75, 1.0, 1.5]
    dev = di.multi_scale_predict(m, x.cuda(), scales, nc, flip=True)
    assert dev.shape == (nc, 97, 129) and torch.isfinite(dev).all()
    with torch.no_grad():
        ref = oi.multi_scale_predict(lambda t: m(t.cuda()), x, scales, nc, torch.device("cuda"), flip=True)
    e = rel(dev.cpu().double().numpy(), ref)
    agree = float((di.predict_labels(dev).cpu().numpy() == ref.argmax(0)).mean())
    msg = f"[inference engine-model] multi-scale+flip scores relerr {e:.2e} label agreement {agree:.4f}"
    print(msg)
    with open(os.path.join(gpu_out_dir, "model_parity.txt"), "a") as f:
        f.write(msg + "\n")
    assert e < 5e-2 and agree > 0.9, msg  # random-init logits have small margins; the scores are the sharp check
    sl = di.sliding_predict(m, x.cuda(), nc, flip=True)
    assert sl.shape == (nc, 97, 129) and torch.isfinite(sl).all()
