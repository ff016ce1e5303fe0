"""Learning-rate schedules of the reference (utils/lr_scheduler.py:4-59) for the fused train step.

The reference's `Poly` / `OneCycle` are `torch.optim.lr_scheduler._LRScheduler` subclasses; with the plugin surface
(`torch.optim.SGD`) they are used unchanged.  `FusedTrainStep` has no `torch.optim` object: its per-parameter rates live in
one device vector and both schedules are a scalar factor on the base rates (OneCycle's `low_lr = lr/div` and
`final_lr = lr/(div*1e4)` are proportional to `lr`), so a schedule here is the same iteration arithmetic producing
`(factor, momentum)` and one `stepper.set_lr_scale(factor)` — no host synchronisation, graph-replay safe (the rates are
read from device memory by the SGD kernel).

Call protocol = the reference's (`trainer.py:52`): construct once (which performs the initial step, as `_LRScheduler`'s
constructor does), then `step(epoch=epoch-1)` before every iteration.  The iteration counter quirk of the reference
(`cur_iter %= iters_per_epoch; cur_iter += 1` *after* T is formed) is reproduced, so the factor sequence is identical.
"""
import math


class _FusedSchedule:
    def __init__(self, stepper, num_epochs, iters_per_epoch=0, last_epoch=-1):
        self.stepper = stepper
        self.iters_per_epoch = iters_per_epoch
        self.cur_iter = 0
        self.N = num_epochs * iters_per_epoch
        self.last_epoch = last_epoch
        self.factor = 1.0
        self.step()  # _LRScheduler.__init__ performs the initial step

    def _advance(self):
        T = self.last_epoch * self.iters_per_epoch + self.cur_iter
        self.cur_iter %= self.iters_per_epoch
        self.cur_iter += 1
        return T

    def get_factor(self):  # pragma: no cover
        raise NotImplementedError

    def step(self, epoch=None):
        """`epoch=None` advances the epoch counter by one (what `_LRScheduler.step()` does), otherwise sets it."""
        self.last_epoch = self.last_epoch + 1 if epoch is None else epoch
        self.factor = self.get_factor()
        self.stepper.set_lr_scale(self.factor)
        return self.factor

    def get_last_lr(self, base_lrs):
        return [b * self.factor for b in base_lrs]


class Poly(_FusedSchedule):
    """utils/lr_scheduler.py:4-21 — factor = (1 - T/N)^0.9, linear warm-up over `warmup_epochs`."""

    def __init__(self, stepper, num_epochs, iters_per_epoch=0, warmup_epochs=0, last_epoch=-1):
        self.warmup_iters = warmup_epochs * iters_per_epoch
        super().__init__(stepper, num_epochs, iters_per_epoch, last_epoch)

    def get_factor(self):
        T = self._advance()
        factor = pow((1 - 1.0 * T / self.N), 0.9)
        if self.warmup_iters > 0 and T < self.warmup_iters:
            factor = 1.0 * T / self.warmup_iters
        return factor


class OneCycle(_FusedSchedule):
    """utils/lr_scheduler.py:24-59 — cosine from lr/div up to lr over `phase1` of the run (momentum 0.95 -> 0.85), then
    cosine down to lr/(div*1e4) (momentum back to 0.95).  Also sets `stepper.momentum`."""

    def __init__(self, stepper, num_epochs, iters_per_epoch=0, last_epoch=-1, momentums=(0.85, 0.95), div_factor=25, phase1=0.3):
        self.N = num_epochs * iters_per_epoch
        self.phase1_iters = int(self.N * phase1)
        self.phase2_iters = self.N - self.phase1_iters
        self.momentums = momentums
        self.mom_diff = momentums[1] - momentums[0]
        self.low = 1.0 / div_factor
        self.final = 1.0 / (div_factor * 1e4)
        super().__init__(stepper, num_epochs, iters_per_epoch, last_epoch)

    def get_factor(self):
        T = self._advance()
        if T <= self.phase1_iters:
            c = (1 + math.cos(math.pi * T / self.phase1_iters)) / 2
            self.stepper.momentum = self.momentums[0] + self.mom_diff * c
            return 1.0 - (1.0 - self.low) * c
        T -= self.phase1_iters
        c = (1 + math.cos(math.pi * T / self.phase2_iters)) / 2
        self.stepper.momentum = self.momentums[1] - self.mom_diff * c
        return self.final + (1.0 - self.final) * c
