"""`seg_b200.optim.SGD` — torch.optim.SGD (as base/base_trainer.py:57 builds it: momentum, weight decay, per-group learning
rates, no dampening / nesterov) with the update done by ONE multi-tensor kernel launch per parameter group
(`seg_sgd_step_dev`) instead of torch's chain of foreach passes.

It IS a torch.optim.SGD: `param_groups`, `state_dict()` / `load_state_dict()` (state = `momentum_buffer` per parameter),
`zero_grad()` and the LR schedulers (which write `group['lr']` / `group['momentum']`) behave identically, so the
reference's `Trainer`, checkpoints and `utils/lr_scheduler.py` work unchanged.  Arithmetic per element, in the order of
torch.optim.sgd:  d = g + wd*p;  buf = momentum*buf + d  (first step: buf = d — the same thing with a zero buffer);
p -= lr*buf.   CUDA fp32 parameters only; anything else raises (no fallback).
"""
import numpy as np
import torch

from . import lib


class SGD(torch.optim.SGD):
    def __init__(self, params, lr=1e-3, momentum=0, dampening=0, weight_decay=0, nesterov=False, **kw):
        if dampening != 0 or nesterov or kw.get("maximize", False):
            raise NotImplementedError("seg_b200.optim.SGD: dampening / nesterov / maximize are not built (the reference's configs use none)")
        for k in ("foreach", "fused", "differentiable"):
            kw.pop(k, None)
        super().__init__(params, lr=lr, momentum=momentum, dampening=0, weight_decay=weight_decay, nesterov=False, **kw)
        self._tables = {}

    def _table(self, gi, params):
        """Device arrays of pointers / sizes for one group, rebuilt (one small H2D copy) when a tensor's storage changes —
        e.g. every step when autograd hands out freshly allocated gradients."""
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["momentum_buffer"].data_ptr()) for p in params)
        t = self._tables.get(gi)
        if t is None or t["key"] != key:
            dev = params[0].device
            n = len(params)
            host = np.empty((4, n), dtype=np.int64)
            host[:3] = np.asarray(key, dtype=np.int64).T
            host[3] = [p.numel() for p in params]
            tab = torch.from_numpy(host).to(dev)
            keep = t if (t is not None and t["lr"].numel() == n and t["lr"].device == dev) else None
            t = {"key": key, "tab": tab, "p": tab[0], "g": tab[1], "m": tab[2], "n": tab[3],
                 "lr": keep["lr"] if keep else torch.empty(n, dtype=torch.float32, device=dev),
                 "hyper": keep["hyper"] if keep else torch.empty(2, dtype=torch.float32, device=dev),
                 "lr_val": keep["lr_val"] if keep else None, "hyper_val": keep["hyper_val"] if keep else None}
            self._tables[gi] = t
        return t

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            for p in params:
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous() and p.grad.dtype == torch.float32):
                    raise RuntimeError("seg_b200.optim.SGD updates contiguous CUDA fp32 parameters only; there is no fallback")
                st = self.state[p]
                if st.get("momentum_buffer") is None:
                    st["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            t = self._table(gi, params)
            lr, hyper = float(group["lr"]), (float(group["momentum"]), float(group["weight_decay"]))
            if t["lr_val"] != lr:
                t["lr"].fill_(lr)
                t["lr_val"] = lr
            if t["hyper_val"] != hyper:
                t["hyper"].copy_(torch.tensor(hyper, dtype=torch.float32))
                t["hyper_val"] = hyper
            lib.call("seg_sgd_step_dev", t["p"].data_ptr(), t["g"].data_ptr(), t["m"].data_ptr(), t["n"].data_ptr(),
                     t["lr"].data_ptr(), len(params), t["hyper"].data_ptr(), 0, 1.0)
            # the kernel writes through raw pointers: tell autograd (and every cache keyed on `_version`, e.g. the engine's
            # packed bf16 weights, engine.ConvSpec.packed) that these tensors changed — exactly what the in-place ops of
            # torch.optim.SGD.step do
            torch.autograd.graph.increment_version(params)
        return loss


_STOCK_SGD = torch.optim.SGD


def make_sgd(params, *args, **kw):
    """What `seg_b200.launch` installs as `torch.optim.SGD` for an unmodified train.py (base/base_trainer.py:57 resolves
    the optimiser by name from `torch.optim`): the multi-tensor-kernel SGD when every parameter is a CUDA fp32 tensor and
    the options are the ones it implements; otherwise the stock class, untouched — the reference's other models (e.g. a
    CPU run of UNet) are not this engine's business."""
    groups = [dict(g, params=list(g["params"])) if isinstance(g, dict) else g for g in params]
    flat = [p for g in groups for p in (g["params"] if isinstance(g, dict) else [g])]
    plain = not (kw.get("nesterov", False) or kw.get("dampening", 0) or kw.get("maximize", False)) and len(args) <= 2
    if plain and flat and all(isinstance(p, torch.Tensor) and p.is_cuda and p.dtype == torch.float32 for p in flat):
        return SGD(groups, *args, **kw)
    return _STOCK_SGD(groups, *args, **kw)


def install():
    torch.optim.SGD = make_sgd
