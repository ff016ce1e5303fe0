"""`utils.metrics` overlay: trainer.py:7 imports eval_metrics / AverageMeter from here; the per-step metric pass becomes one
device kernel + one small copy (seg_b200.metrics).  The helper functions of the reference module stay reachable."""
import importlib.util
import os

from . import REFERENCE_UTILS

if REFERENCE_UTILS is not None:
    _spec = importlib.util.spec_from_file_location("_reference_utils_metrics", os.path.join(REFERENCE_UTILS, "metrics.py"))
    _ref = importlib.util.module_from_spec(_spec)
    _spec.loader.exec_module(_ref)
    for _n in dir(_ref):
        if not _n.startswith("_"):
            globals()[_n] = getattr(_ref, _n)

from seg_b200.metrics import AverageMeter, eval_metrics  # noqa: E402,F401
