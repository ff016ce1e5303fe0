"""seg_b200 — B200-native forward/backward engine for the data-parallel hot path of yassouali/pytorch-segmentation.

Host side mirrors the reference's plugin surface (models / losses registries); all arithmetic runs in the hand-written
sm_100a kernels of ``libseg_b200.so`` (C ABI: include/seg_b200.h).  No CPU fallback.
"""
from . import lib  # noqa: F401
from .losses import CrossEntropyLoss2d  # noqa: F401
from .nets import DeepLab, PSPNet  # noqa: F401

__all__ = ["DeepLab", "PSPNet", "CrossEntropyLoss2d", "lib"]
