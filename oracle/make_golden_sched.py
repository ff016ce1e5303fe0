#!/usr/bin/env python
"""Golden vectors for utils/lr_scheduler.py (Poly, OneCycle), produced by importing the REFERENCE (read-only,
/root/reference) in this container and driving it exactly as the trainer does (base_trainer.py:58, trainer.py:52:
construct, then `step(epoch=epoch-1)` before every iteration).  Writes tests/golden/lr_sched.npz.
Run:  python oracle/make_golden_sched.py"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SEG_REFERENCE_ROOT", "/root/reference")


def drive(cls, epochs, iters, **kw):
    dec, bb = [torch.nn.Parameter(torch.zeros(3))], [torch.nn.Parameter(torch.zeros(2))]
    opt = torch.optim.SGD([{"params": dec}, {"params": bb, "lr": 0.001}], lr=0.01, momentum=0.9, weight_decay=1e-4)
    sched = cls(opt, epochs, iters, **kw)
    lrs, moms = [], []
    for epoch in range(1, epochs + 1):
        for _ in range(iters):
            sched.step(epoch=epoch - 1)
            lrs.append([g["lr"] for g in opt.param_groups])
            moms.append([g["momentum"] for g in opt.param_groups])
    return np.asarray(lrs, dtype=np.float64), np.asarray(moms, dtype=np.float64)


def main():
    warnings.filterwarnings("ignore")
    sys.path.insert(0, REF)
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_lr_scheduler", os.path.join(REF, "utils", "lr_scheduler.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rec = {}
    for tag, cls, epochs, iters, kw in (("poly", mod.Poly, 4, 7, {}), ("poly_warm", mod.Poly, 5, 6, {"warmup_epochs": 2}),
                                         ("onecycle", mod.OneCycle, 4, 9, {}), ("onecycle_p5", mod.OneCycle, 3, 10, {"phase1": 0.5, "div_factor": 10})):
        lrs, moms = drive(cls, epochs, iters, **kw)
        rec[f"{tag}/lrs"], rec[f"{tag}/moms"] = lrs, moms
        rec[f"{tag}/cfg"] = np.asarray([epochs, iters], dtype=np.int64)
        print(tag, lrs[0], lrs[-1], moms[0], moms[-1])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "lr_sched.npz"), **rec)


if __name__ == "__main__":
    main()
