"""seg_b200.lr_scheduler (the fused step's Poly / OneCycle) against golden vectors produced by the reference's own
utils/lr_scheduler.py driven the way trainer.py:52 drives it (oracle/make_golden_sched.py)."""
import os

import numpy as np
import pytest

from seg_b200 import lr_scheduler as sched

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lr_sched.npz")


class FakeStepper:
    """The two things a schedule touches on FusedTrainStep: `set_lr_scale(factor)` and the `momentum` attribute."""

    def __init__(self, base_lrs=(0.01, 0.001), momentum=0.9):
        self.base_lrs, self.momentum, self.scale = base_lrs, momentum, 1.0

    def set_lr_scale(self, scale):
        self.scale = float(scale)


@pytest.mark.parametrize("tag,cls,kw", [("poly", sched.Poly, {}), ("poly_warm", sched.Poly, {"warmup_epochs": 2}),
                                        ("onecycle", sched.OneCycle, {}),
                                        ("onecycle_p5", sched.OneCycle, {"phase1": 0.5, "div_factor": 10})])
def test_schedule_matches_reference(tag, cls, kw):
    g = np.load(GOLD)
    epochs, iters = (int(v) for v in g[f"{tag}/cfg"])
    st = FakeStepper()
    s = cls(st, epochs, iters, **kw)
    lrs, moms = [], []
    for epoch in range(1, epochs + 1):
        for _ in range(iters):
            s.step(epoch=epoch - 1)
            lrs.append([b * st.scale for b in st.base_lrs])
            moms.append([st.momentum, st.momentum])
    np.testing.assert_allclose(np.asarray(lrs), g[f"{tag}/lrs"], rtol=1e-12, atol=1e-18)
    np.testing.assert_allclose(np.asarray(moms), g[f"{tag}/moms"], rtol=1e-12, atol=0)
