#!/usr/bin/env python
"""CPU experiment behind DESIGN.md §4 "run-to-run determinism": which layer turns a last-bit difference of the BatchNorm
statistics into percent-level logit differences at batch 2?

The engine's host logic runs here on the ATen emulation of the kernels (tests/cpu_emulation.py, bf16 activation storage
like the device).  Run A is the plain forward; run B perturbs every BN statistics vector by one fp32 ulp-sized relative
factor with a random sign (what a different atomic-add order does on the device).  Every BN layer's output is recorded
in both runs; the script prints, per layer, the largest relative difference and the number of samples per channel the
layer normalises over, so the amplifier can be read off.

    python tools/chaos_probe_cpu.py            # DeepLab/ResNet-14, 2 x 3 x 65 x 65
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-segmentation_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

import cpu_emulation as emu  # noqa: E402
from seg_b200 import engine, nets  # noqa: E402
from seg_b200 import losses as plosses  # noqa: E402
from oracle import synth, weights  # noqa: E402


def main():
    torch.manual_seed(0)
    torch.set_num_threads(4)
    for mod in (engine, nets, plosses):
        mod.ops = emu
    nets._EngineModel._check_input = lambda self, x: None
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    sd = weights.deeplab_resnet_state_dict(7, "resnet14", seed=21, randomize_bn=True)
    x, _ = synth.make_batch(batch, 65, 65, 7, 255, seed=9100)

    real_apply = emu.bn_apply_train
    state = {"perturb": False, "log": None, "gen": None}

    def bn_apply_train(xin, stats, count, *a, **k):
        if state["perturb"]:
            sign = (torch.rand(stats.shape, generator=state["gen"]) > 0.5).float() * 2 - 1
            stats = stats * (1 + sign * 6e-8)  # one fp32 ulp, random direction per channel and moment
        out, save = real_apply(xin, stats, count, *a, **k)
        state["log"].append((tuple(xin.shape), float(count), out.float().clone()))
        return out, save

    emu.bn_apply_train = bn_apply_train
    runs = []
    for perturb in (False, True):
        m = nets.DeepLab(7, backbone="resnet14", output_stride=16)
        m.load_state_dict(sd, strict=True)
        m.engine_dropout = False
        m.train()
        state.update(perturb=perturb, log=[], gen=torch.Generator().manual_seed(5))
        with torch.no_grad():
            out = m(x)
        runs.append((out.float(), state["log"]))
    (oa, la), (ob, lb) = runs
    print(f"logits: max relative difference {((oa - ob).abs().max() / oa.abs().max()).item():.3e}  (batch {batch})")
    print(f"{'#':>3} {'BN input [N,H,W,C]':>24} {'samples/channel':>16} {'max rel diff of the BN output':>32} {'channels with diff > 1e-2':>28}")
    for i, ((shape, count, a), (_, _, b)) in enumerate(zip(la, lb)):
        d = (a - b).abs()
        rel = (d.max() / (a.abs().max() + 1e-12)).item()
        per_ch = d.reshape(-1, shape[-1]).max(0).values / (a.abs().max() + 1e-12)
        print(f"{i:3d} {str(shape):>24} {int(count):16d} {rel:32.3e} {int((per_ch > 1e-2).sum()):28d}")


if __name__ == "__main__":
    main()
