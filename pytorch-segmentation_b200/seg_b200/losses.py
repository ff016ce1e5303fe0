"""Drop-in losses for the reference's loss registry (`getattr(losses, config['loss'])(ignore_index=...)`, train.py:30).

    CrossEntropyLoss2d(weight=None, ignore_index=255, reduction='mean')   — utils/losses.py:24-31

forward(output fp32 [B,C,H,W], target int64 [B,H,W]) -> 0-dim tensor with autograd, computed by the sm_100a kernels
(`seg_ce_nchw_fwd/bwd`); `.item()` works as the trainer expects (trainer.py:72,81).  CUDA tensors only.
"""
import torch
import torch.nn as nn

from . import ops


class _CEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, ignore_index):
        logits = logits.contiguous().float()
        target = target.contiguous()
        loss, accum = ops.ce_nchw_fwd(logits, target, ignore_index)
        ctx.save_for_backward(logits, target, accum)
        ctx.ignore_index = ignore_index
        return loss

    @staticmethod
    def backward(ctx, gout):
        logits, target, accum = ctx.saved_tensors
        g = gout.detach().reshape(1).float().contiguous()
        return ops.ce_nchw_bwd(logits, target, ctx.ignore_index, accum, gscale=g), None, None


class CrossEntropyLoss2d(nn.Module):
    def __init__(self, weight=None, ignore_index=255, reduction="mean"):
        super().__init__()
        if weight is not None or reduction != "mean":
            raise NotImplementedError("seg_b200.CrossEntropyLoss2d: only weight=None, reduction='mean' (the configs' setting)")
        self.ignore_index = ignore_index

    def forward(self, output, target):
        if not output.is_cuda:
            raise RuntimeError("seg_b200 losses run on a B200 only; there is no CPU fallback")
        return _CEFn.apply(output, target, self.ignore_index)
