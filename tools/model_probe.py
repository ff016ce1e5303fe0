#!/usr/bin/env python
"""Throughput probe of the secondary configs through FusedTrainStep (CE loss), CUDA-event timed.
    python tools/model_probe.py --config c2|c4|c5 [--batch B] [--steps K]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pytorch-segmentation_b200"))
import torch  # noqa: E402
import seg_b200  # noqa: E402
from seg_b200.train import FusedTrainStep  # noqa: E402
from oracle import synth  # noqa: E402

CFG = {
    "c2": dict(name="PSPNet/ResNet50 473x473 21cls (CE + 0.4 aux CE)", size=473, nc=21, ignore=255, batch=16, gflop=1071.78,
               build=lambda: seg_b200.PSPNet(21, backbone="resnet50")),
    "c4": dict(name="DeepLabV3+/Xception 769x769 19cls (CE)", size=769, nc=19, ignore=255, batch=8, gflop=1150.90,
               build=lambda: seg_b200.DeepLab(19, backbone="xception")),
    "c5": dict(name="UperNet/ResNet101 512x512 150cls (CE; the reference config uses Lovasz)", size=512, nc=150, ignore=-1, batch=8,
               gflop=1123.96, build=lambda: seg_b200.UperNet(150, backbone="resnet101")),
}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--graph", type=int, default=1)
    a = ap.parse_args()
    c = CFG[a.config]
    B = a.batch or c["batch"]
    torch.manual_seed(0)
    m = c["build"]().cuda().train()
    x, y = synth.make_batch(B, c["size"], c["size"], c["nc"], c["ignore"], seed=1234)
    x, y = x.cuda(), y.cuda()
    st = FusedTrainStep(m, ignore_index=c["ignore"], cuda_graph=bool(a.graph))
    for _ in range(3):
        st.step(x, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = st.step(x, y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(json.dumps({"config": c["name"], "batch": B, "ms_per_step": ms, "images_per_sec": B / ms * 1e3,
                      "conv_tflops": B / ms * 1e3 * c["gflop"] / 1e3, "loss": float(loss.item()), "cuda_graph": bool(a.graph),
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))
