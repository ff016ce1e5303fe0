"""Forward/backward engine: a tape of hand-written-kernel ops over NHWC bf16 activations.

The reference runs the hot path through torch.autograd over ATen ops (trainer.py:56,70).  Here the model's forward
is a straight-line sequence of C-ABI kernel calls recorded on a `Tape`; `Tape.backward()` replays the recorded
closures in reverse (dgrad / wgrad / BN-backward / resize-backward kernels).  torch is used for device memory only.

Conventions
  * `Act` = an activation: `.t` NHWC bf16 tensor (possibly a channel slice of a concat buffer), `.grad` its
    gradient buffer (allocated on first write; later writers accumulate with beta = 1).
  * BatchNorm (training) = statistics fused into the producing conv's epilogue -> `bn_finalize` -> one fused
    apply(+residual +ReLU +dropout) pass.  Backward = one reduce pass + one apply pass.
  * concatenation never copies: producers write into channel slices of one buffer (`Tape.concat`).
"""
import torch

from . import ops
from .lib import IMPL_AUTO

BN_EPS = 1e-5
BN_MOM = 0.1
FUSED_BWD_MAX_BYTES = 24 << 20  # BN backward: one cooperative launch up to this activation size (bf16 bytes), two above
ACT_DTYPE = torch.bfloat16  # storage type of activations / activation gradients (the kernels are bf16-only;
                            # tests/test_engine_cpu_emulated.py flips this to fp32 together with the ATen emulation)


class Act:
    __slots__ = ("t", "grad", "needs_grad", "_written")

    def __init__(self, t, needs_grad=True):
        self.t = t
        self.grad = None
        self.needs_grad = needs_grad
        self._written = False

    @property
    def shape(self):
        return self.t.shape

    def grad_target(self):
        """(buffer, beta): beta = 0 for the first writer of this gradient, 1 afterwards."""
        if self.grad is None:
            self.grad = torch.empty(self.t.shape, dtype=ACT_DTYPE, device=self.t.device)
        if not self._written:
            self._written = True
            return self.grad, 0.0
        return self.grad, 1.0


class ConvSpec:
    """Static description of one nn.Conv2d holder + its packed bf16 weight cache."""

    def __init__(self, name, module, explicit_im2col=False):
        self.name = name
        self.m = module
        w = module.weight
        self.K, self.C, self.R, self.S = w.shape
        assert module.stride[0] == module.stride[1] and module.padding[0] == module.padding[1]
        assert module.dilation[0] == module.dilation[1] and module.groups == 1
        self.explicit = explicit_im2col or (self.C % 8 != 0)
        self.kpad = (self.R * self.S * self.C + 7) // 8 * 8 if self.explicit else None
        self._packed = None
        self._version = None
        self.always_repack = False  # set while a CUDA graph of the train step is captured / replayed

    @property
    def stride(self):
        return self.m.stride[0]

    @property
    def pad(self):
        return self.m.padding[0]

    @property
    def dil(self):
        return self.m.dilation[0]

    def packed_shape(self):
        return (1, self.K, self.kpad) if self.explicit else (self.R * self.S, self.K, self.C)

    def packed(self):
        w = self.m.weight
        key = (w._version, w.data_ptr())
        if self._packed is None or self._version != key or self.always_repack:
            if self.explicit:
                # column order (r, s, c): OIHW -> O,(H,W,I) as a 1x1 weight over Kpad "channels"
                w2 = w.detach().permute(0, 2, 3, 1).reshape(self.K, self.R * self.S * self.C, 1, 1).contiguous()
                self._packed = ops.pack_weight(w2, cpad=self.kpad)
            else:
                self._packed = ops.pack_weight(w.detach())
            self._version = key
        return self._packed


class DwSpec:
    """Depthwise 3x3 conv holder (nn.Conv2d with groups == channels) + packed fp32 [9][C] weight cache."""

    def __init__(self, name, module):
        self.name, self.m = name, module
        assert module.groups == module.in_channels == module.out_channels and module.kernel_size == (3, 3)
        self.C = module.in_channels
        self._packed, self._version = None, None
        self.always_repack = False

    stride = property(lambda self: self.m.stride[0])
    pad = property(lambda self: self.m.padding[0])
    dil = property(lambda self: self.m.dilation[0])

    def packed(self):
        w = self.m.weight
        key = (w._version, w.data_ptr())
        if self._packed is None or self._version != key or self.always_repack:
            self._packed = ops.dw_pack_weight(w.detach())
            self._version = key
        return self._packed


class Tape:
    def __init__(self, training, record=None, grads=None, impl=IMPL_AUTO, dropout=True, seed=0, sync=None, clamp_eps=False,
                 step_ctr=None, arena_floats=0, owner=None):
        self.training = training                                  # module.training semantics (batch stats, dropout)
        self.record = training if record is None else record      # record backward closures
        self.back = []
        self.back_params = {}  # index into self.back -> parameters whose gradient that closure writes (bucketed all-reduce)
        self.grads = grads if grads is not None else {}  # Parameter -> fp32 grad tensor (param layout)
        self.touched = set()  # parameters whose gradient this tape's backward wrote (pre-bound `grads` hide that)
        self.impl = impl
        self.dropout = dropout
        self.seed = seed
        self.sync = sync  # object with .allreduce_(fp32 vector) and .world ; None = local BN
        self.step_ctr = step_ctr  # device int64 step counter mixed into dropout seeds / SyncBN epochs (graph-replay safe)
        # FusedTrainStep hands in persistent buffers: bf16 packed weights (one batched pack launch per step) and fp32
        # packed weight-gradient accumulators (one batched unpack launch per step).  Empty for the autograd/plugin path.
        self.packed_override = {}
        self.dw_buffers = {}
        self.clamp_eps = clamp_eps
        self._drop_ctr = 0
        self._pushed = set()  # statistics vectors whose producing conv already pushed them to the SyncBN peers
        self.bn_modules = []
        # zero-initialised fp32 arena: BN-statistics accumulators and reduction slots of the whole step come out of ONE
        # zero fill instead of one fill per layer (arena_floats = what the previous step used; grows in 1M-float chunks)
        self.owner = owner  # the model: remembers how much arena a step needs
        self._arena_hint = int(arena_floats)
        self._arena = None
        self._arena_off = 0
        self.arena_used = 0

    # ------------------------------------------------------------------ helpers
    def sync_active(self):
        """SyncBN exchange on: more than one rank, or a loopback group forced on for the single-GPU protocol test."""
        s = self.sync
        return s is not None and (s.world > 1 or getattr(s, "force", False))

    def sync_fused(self):
        return self.sync_active() and getattr(self.sync, "fused", False)

    def zalloc(self, n, device):
        """n zeroed floats (128-byte aligned) from the step's arena."""
        n_al = (int(n) + 31) // 32 * 32
        if self._arena is None or self._arena_off + n_al > self._arena.numel() or self._arena.device != device:
            size = max(self._arena_hint if self._arena is None else 0, n_al, 1 << 20)
            self._arena = torch.zeros(size, dtype=torch.float32, device=device)
            self._arena_off = 0
        v = self._arena[self._arena_off:self._arena_off + int(n)]
        self._arena_off += n_al
        self.arena_used += n_al
        return v

    def zalloc64(self, n, device):
        """n zeroed fp64 words from the step's arena (the arena hands out 128-byte aligned float ranges)."""
        return self.zalloc(2 * int(n), device).view(torch.float64)

    def _param_grad(self, p, value_fn):
        """Write (or accumulate into) the fp32 gradient of parameter p.  value_fn(out, beta) fills it."""
        if not p.requires_grad:
            return
        self.touched.add(p)
        if p in self.grads:
            value_fn(self.grads[p], 1.0)
        else:
            g = torch.empty(p.shape, dtype=torch.float32, device=p.device)
            value_fn(g, 0.0)
            self.grads[p] = g

    def _push_back(self, fn, params=()):
        if params:
            self.back_params[len(self.back)] = tuple(p for p in params if p is not None)
        self.back.append(fn)

    def backward(self, after=None):
        """Replay the recorded closures in reverse.  after(i): called when closure i (and everything recorded after it) has
        run — the fused train step uses it to launch a gradient bucket's all-reduce as soon as the bucket is complete."""
        for i in range(len(self.back) - 1, -1, -1):
            self.back[i]()
            if after is not None:
                after(i)
        if self.owner is not None and self.arena_used:
            self.owner._arena_floats = self.arena_used
        self.back = []

    # ------------------------------------------------------------------ conv
    def conv(self, x, spec, out=None, out_dtype=None, want_stats=False):
        """x: Act (NHWC bf16) — or, for an explicit-im2col conv, a raw NCHW fp32 tensor (the network input)."""
        wp = self.packed_override.get(spec)
        if wp is None:
            wp = spec.packed()
        bias = spec.m.bias
        stats = None
        if out_dtype is None:
            out_dtype = ACT_DTYPE
        want = want_stats and self.training
        sync = tk = None
        if want:
            stats = self.zalloc64(2 * spec.K, wp.device)  # fp64 accumulators: the conv epilogue adds its column sums
            if self.sync_fused():  # SyncBN: the last CTA pushes the totals to the peers (no exchange launch)
                sync, tk = self.sync, self.zalloc(1, wp.device)
        if spec.explicit:
            nchw = not isinstance(x, Act)
            src = x if nchw else x.t
            col = ops.im2col(src, spec.R, spec.S, spec.stride, spec.pad, spec.dil, spec.kpad, nchw_f32=nchw)
            y = ops.conv2d_fwd(col, wp, spec.K, 1, 1, out=out, out_dtype=out_dtype, bias=bias.detach() if bias is not None else None,
                               stats=stats, impl=self.impl, sync=sync, sync_ticket=tk)
            xin, geo = col, (1, 1, 1, 0, 1)
        else:
            y = ops.conv2d_fwd(x.t, wp, spec.K, spec.R, spec.S, spec.stride, spec.pad, spec.dil, out=out, out_dtype=out_dtype,
                               bias=bias.detach() if bias is not None else None, stats=stats, impl=self.impl, sync=sync, sync_ticket=tk)
            xin, geo = x.t, (spec.R, spec.S, spec.stride, spec.pad, spec.dil)
        ya = Act(y)
        if sync is not None:
            self._pushed.add(stats.data_ptr())
        if self.record:
            def bwd():
                dy = ya.grad
                if dy is None:
                    return
                R, S, stride, pad, dil = geo
                if spec.m.weight.requires_grad and spec in self.dw_buffers:
                    # accumulate into the trainer's persistent packed-gradient buffer; unpacked once per step, batched
                    ops.conv2d_wgrad(dy, xin, R, S, stride, pad, dil, out=self.dw_buffers[spec], impl=self.impl)
                    self.touched.add(spec.m.weight)
                elif spec.m.weight.requires_grad:
                    dwp = ops.conv2d_wgrad(dy, xin, R, S, stride, pad, dil, impl=self.impl)
                    if spec.explicit:
                        def fill(g, beta):
                            # packed [1][K][Kpad] -> [K][R,S,C] -> OIHW
                            full = dwp[0, :, : spec.R * spec.S * spec.C].reshape(spec.K, spec.R, spec.S, spec.C).permute(0, 3, 1, 2)
                            if beta:
                                g.add_(full)
                            else:
                                g.copy_(full)
                    else:
                        def fill(g, beta):
                            ops.unpack_wgrad(dwp, tuple(spec.m.weight.shape), beta=beta, out=g)
                    self._param_grad(spec.m.weight, fill)
                if bias is not None and bias.requires_grad:
                    cpad = ops.ld(dy)
                    wide = dy if dy.shape[-1] % 8 == 0 else dy.as_strided(dy.shape[:-1] + (cpad,), dy.stride(), dy.storage_offset())
                    s = ops.bn_stats(wide)[: spec.K].float()  # column sums of dY (fp64 accumulation)
                    self._param_grad(bias, lambda g, beta: g.add_(s) if beta else g.copy_(s))
                if (not spec.explicit) and isinstance(x, Act) and x.needs_grad:
                    gx, beta = x.grad_target()
                    ops.conv2d_dgrad(dy, wp, tuple(x.t.shape), R, S, stride, pad, dil, out=gx, beta=beta, impl=self.impl)
                ya.grad = None
            self._push_back(bwd, (spec.m.weight, bias))
        return ya, stats

    def dwconv(self, x, spec, want_stats=False):
        """Depthwise 3x3 (SeparableConv2d.conv1).  Returns (raw output Act, BN statistics or None)."""
        w9 = spec.packed()
        stats = self.zalloc64(2 * spec.C, w9.device) if (want_stats and self.training) else None
        sync = self.sync if (stats is not None and self.sync_fused()) else None
        y = ops.dwconv_fwd(x.t, w9, spec.stride, spec.pad, spec.dil, stats=stats, sync=sync,
                           sync_ticket=self.zalloc(1, w9.device) if sync is not None else None)
        if sync is not None:
            self._pushed.add(stats.data_ptr())
        ya = Act(y)
        if self.record:
            def bwd():
                dy = ya.grad
                if dy is None:
                    return
                if spec.m.weight.requires_grad:
                    g9 = ops.dwconv_bwd_weight(dy, x.t, spec.stride, spec.pad, spec.dil)
                    self._param_grad(spec.m.weight, lambda g, beta: ops.dw_unpack_wgrad(g9, g, beta))
                if x.needs_grad:
                    gx, beta = x.grad_target()
                    ops.dwconv_bwd_data(dy, w9, tuple(x.t.shape), spec.stride, spec.pad, spec.dil, out=gx, beta=beta)
                ya.grad = None
            self._push_back(bwd, (spec.m.weight,))
        return ya, stats

    def relu(self, x):
        y = ops.relu_fwd(x.t)
        ya = Act(y)
        if self.record:
            def bwd():
                if ya.grad is None or not x.needs_grad:
                    return
                gx, beta = x.grad_target()
                ops.relu_bwd(ya.grad, y, gx, beta)
                ya.grad = None
            self.back.append(bwd)
        return ya

    # ------------------------------------------------------------------ batch norm (+ residual, ReLU, dropout)
    def bn_act(self, y, bn, stats=None, relu=True, res=None, out=None, drop_p=0.0, drop_channelwise=False):
        """y: Act holding the raw conv output.  Returns the activated Act.  drop_channelwise: nn.Dropout2d semantics (one
        draw per image and channel, models/pspnet.py:22,68) instead of nn.Dropout's per-element draws."""
        drop_hw = y.t.shape[1] * y.t.shape[2] if drop_channelwise else 0
        C = y.t.shape[-1]
        count_local = ops.rows(y.t)
        use_batch_stats = self.training and bn.training
        if self.training and not (self.dropout):
            drop_p = 0.0
        if not self.training:
            drop_p = 0.0
        seed = 0
        if drop_p > 0.0:
            self._drop_ctr += 1
            seed = (self.seed * 1000003 + self._drop_ctr * 7919) & 0x7FFFFFFFFFFFFFFF  # + step counter on the device
        if use_batch_stats:
            if stats is None:
                sync0 = self.sync if self.sync_fused() else None
                stats = ops.bn_stats(y.t, sync=sync0)
                if sync0 is not None:
                    self._pushed.add(stats.data_ptr())
            count = count_local
            in_kernel = None
            if self.sync_active():
                if stats.data_ptr() in self._pushed:
                    if getattr(self.sync, "mode", 0) == 0:
                        in_kernel = self.sync  # bn_apply waits for the world's flags and adds every rank's sums itself
                    # mode 1: the producer's last block already did the exchange; `stats` holds the world's totals
                else:  # exchange object without the in-kernel protocol (the gloo stand-in of the CPU tests): sums over ranks
                    self.sync.allreduce_(stats)
                count = count_local * self.sync.world
            # finalize (coefficients, saved mean / 1/std, running statistics) happens inside the apply kernel
            a, save = ops.bn_apply_train(y.t, stats, count, bn.weight.detach(), bn.bias.detach(), bn.eps,
                                         bn.momentum if bn.momentum is not None else BN_MOM,
                                         1 if (self.clamp_eps and self.sync is not None and self.sync.world > 1) else 0,
                                         bn.running_mean, bn.running_var, res=res.t if res is not None else None, out=out,
                                         relu=relu, drop_p=drop_p, seed=seed, step_ctr=self.step_ctr if drop_p > 0.0 else None,
                                         drop_hw=drop_hw, sync=in_kernel,
                                         sync_done=self.zalloc(1, y.t.device) if in_kernel is not None else None)
            self.bn_modules.append(bn)
        else:
            ss, save = ops.bn_eval_scale_shift(bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, bn.eps,
                                               want_save=True)
            count = count_local
            a = ops.bn_apply(y.t, ss, res=res.t if res is not None else None, out=out, relu=relu, drop_p=drop_p, seed=seed,
                             step_ctr=self.step_ctr if drop_p > 0.0 else None, drop_hw=drop_hw)
        aa = Act(a)
        # conv -> BN(batch statistics) -> ReLU with nothing in between: the backward APPLY pass recomputes the ReLU mask
        # from the conv output with the forward's own coefficients instead of re-reading the activation (one stream less).
        # The reduce pass keeps reading the activation: its mask-recomputing variant needs 74 registers (3 blocks/SM
        # instead of 4) and measured slower.
        remask = bool(use_batch_stats and relu and res is None and drop_p == 0.0)
        if self.record:
            def bwd():
                da = aa.grad
                if da is None:
                    return
                want_pg = bn.weight.requires_grad
                acc_pg = want_pg and (bn.weight in self.grads)
                if want_pg:
                    self.touched.update((bn.weight, bn.bias))
                if want_pg and not acc_pg:
                    self.grads[bn.weight] = torch.empty(C, dtype=torch.float32, device=a.device)
                    self.grads[bn.bias] = torch.empty(C, dtype=torch.float32, device=a.device)
                a_mask = None if remask else a
                dy = torch.empty(y.t.shape, dtype=ACT_DTYPE, device=a.device)
                dres, beta_res = (None, 0.0)
                if res is not None and res.needs_grad:
                    dres, beta_res = res.grad_target()
                sync = self.sync if (use_batch_stats and self.sync_active()) else None
                dg = self.grads[bn.weight] if want_pg else None
                db = self.grads[bn.bias] if want_pg else None
                if sync is not None and not getattr(sync, "fused", False):
                    # exchange object without the in-kernel protocol (the gloo stand-in of the CPU tests)
                    sums = ops.bn_bwd_reduce(da, a, y.t, save, relu=relu, drop_p=drop_p, dgamma=dg, dbeta=db, accumulate=acc_pg,
                                             acc=self.zalloc64(ops.bn_bwd_reduce_acc_words(C), a.device))
                    gsums = sums.clone()
                    sync.allreduce_(gsums)
                    ops.bn_bwd_apply(da, a_mask, y.t, save, bn.weight.detach(), gsums, count, relu=relu, drop_p=drop_p, dx=dy,
                                     dres=dres, beta_res=beta_res, beta=bn.bias.detach())
                elif count_local * C * 2 <= FUSED_BWD_MAX_BYTES:
                    # small maps (the operands stay in L2 between the phases): reduce -> grid barrier -> fixed-order cross-block
                    # sum (-> SyncBN exchange) -> apply in ONE cooperative launch.  Frozen BN (freeze_bn): dx = gamma*inv_std*dz,
                    # the sums only feed the parameter gradients
                    ops.bn_bwd_fused(da, a_mask, y.t, save, bn.weight.detach(), count, relu=relu, drop_p=drop_p, dgamma=dg, dbeta=db,
                                     accumulate=acc_pg, dx=dy, dres=dres, beta_res=beta_res, beta=bn.bias.detach(),
                                     zero_sums=not use_batch_stats, tickets=self.zalloc(2, a.device), sync=sync)
                else:
                    # large maps stream from HBM in both passes anyway: two launches at full occupancy; under SyncBN the
                    # reduction's last block pushes the sums and the apply pass waits for the world's
                    sums = ops.bn_bwd_reduce(da, a, y.t, save, relu=relu, drop_p=drop_p, dgamma=dg, dbeta=db, accumulate=acc_pg,
                                             acc=self.zalloc64(ops.bn_bwd_reduce_acc_words(C), a.device), sync=sync)
                    gsums = sums if use_batch_stats else torch.zeros_like(sums)
                    csync = sync if (sync is not None and getattr(sync, "mode", 0) == 0) else None  # mode 1: sums are the world's
                    ops.bn_bwd_apply(da, a_mask, y.t, save, bn.weight.detach(), gsums, count, relu=relu, drop_p=drop_p, dx=dy,
                                     dres=dres, beta_res=beta_res, beta=bn.bias.detach(), sync=csync,
                                     sync_done=self.zalloc(1, a.device) if csync is not None else None)
                y.grad = dy
                aa.grad = None
            self._push_back(bwd, (bn.weight, bn.bias))
        return aa

    # ------------------------------------------------------------------ pooling / resize / concat
    def maxpool(self, x):
        y, idx = ops.maxpool3x3s2_fwd(x.t)
        ya = Act(y)
        if self.record:
            def bwd():
                if ya.grad is None or not x.needs_grad:
                    return
                assert x.grad is None, "maxpool input must have a single consumer"
                x.grad = ops.maxpool3x3s2_bwd(ya.grad, idx, tuple(x.t.shape))
                ya.grad = None
            self.back.append(bwd)
        return ya

    def avgpool(self, x, bins):
        y = ops.adaptive_avgpool_fwd(x.t, bins)
        ya = Act(y)
        if self.record:
            def bwd():
                if ya.grad is None or not x.needs_grad:
                    return
                gx, beta = x.grad_target()
                ops.adaptive_avgpool_bwd(ya.grad, tuple(x.t.shape), bins, dx=gx, beta=beta)
                ya.grad = None
            self.back.append(bwd)
        return ya

    def bilinear(self, x, Ho, Wo, align_corners, out=None):
        y = ops.bilinear_fwd(x.t, Ho, Wo, align_corners, out=out)
        ya = Act(y)
        if self.record:
            def bwd():
                if ya.grad is None or not x.needs_grad:
                    return
                gx, beta = x.grad_target()
                ops.bilinear_bwd(ya.grad, x.t.shape[1], x.t.shape[2], align_corners, dx=gx, beta=beta)
                ya.grad = None
            self.back.append(bwd)
        return ya

    def up_add(self, x, y):
        """up_and_add of the reference's FPN (models/upernet.py:89-90): bilinear(x -> size of y, align_corners=True) + y."""
        Hy, Wy = y.t.shape[1], y.t.shape[2]
        u = ops.bilinear_fwd(x.t, Hy, Wy, True)
        ops.axpby(y.t, u, 1.0)  # u += y
        ua = Act(u)
        if self.record:
            def bwd():
                g = ua.grad
                if g is None:
                    return
                if y.needs_grad:
                    gy, beta = y.grad_target()
                    ops.axpby(g, gy, beta)
                if x.needs_grad:
                    gx, beta = x.grad_target()
                    ops.bilinear_bwd(g, x.t.shape[1], x.t.shape[2], True, dx=gx, beta=beta)
                ua.grad = None
            self.back.append(bwd)
        return ua

    def copy_into(self, x, out):
        """Copy an activation into a concat slice (used when the producer's tensor is also consumed elsewhere)."""
        ops.axpby(x.t, out, 0.0)
        ya = Act(out)
        if self.record:
            def bwd():
                if ya.grad is None or not x.needs_grad:
                    return
                gx, beta = x.grad_target()
                ops.axpby(ya.grad, gx, beta)
                ya.grad = None
            self.back.append(bwd)
        return ya

    def concat(self, N, H, W, channels, device):
        """Allocate one NHWC buffer; returns (whole Act, [slice tensors]).  Producers pass a slice as `out=`; call
        `bind_slices` with the producer Acts so their gradients alias slices of the whole gradient."""
        total = sum(channels)
        buf = torch.empty((N, H, W, total), dtype=ACT_DTYPE, device=device)
        whole = Act(buf)
        slices, off = [], 0
        for c in channels:
            slices.append(buf[..., off:off + c])
            off += c
        return whole, slices

    def bind_slices(self, whole, acts):
        """After the producers ran: make each producer Act's gradient a view of the concat gradient buffer."""
        if not self.record:
            return
        g = torch.empty(whole.t.shape, dtype=ACT_DTYPE, device=whole.t.device)
        whole.grad = g
        off = 0
        for a in acts:
            c = a.t.shape[-1]
            a.grad = g[..., off:off + c]
            off += c
        whole._written = False  # the consumer's dgrad overwrites (beta = 0) the pre-allocated buffer
