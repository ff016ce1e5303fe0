#!/usr/bin/env python
"""Micro-benchmark of the three streaming BatchNorm kernels through the C ABI on one activation shape.

    python tools/bn_micro.py --shape 17424,256 --shape 17424,1024 --res 1
Buffers are rotated (working set > L2) and each kernel is timed with CUDA events over `iters` launches.
Prints us/launch and algorithmic TB/s (bf16 passes over rows x C).
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-segmentation_b200"))
import torch  # noqa: E402
from seg_b200 import lib, ops  # noqa: E402


def timeit(fn, iters):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def run(M, C, res, iters):
    nbuf = max(2, int(400e6 // (M * C * 2 * 4)) + 1)
    mk = lambda: torch.randn(M, C, device="cuda").to(torch.bfloat16)
    xs, dys, outs, rs = [mk() for _ in range(nbuf)], [mk() for _ in range(nbuf)], [mk() for _ in range(nbuf)], [mk() for _ in range(nbuf)]
    dxs = [torch.empty(M, C, device="cuda", dtype=torch.bfloat16) for _ in range(nbuf)]
    gamma, beta = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda")
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    stats = ops.bn_stats(xs[0])
    n = ops.bn_bwd_reduce_acc_words(C) + 1
    zs = torch.zeros(n * (iters + 16), dtype=torch.float64, device="cuda")
    _, save = ops.bn_apply_train(xs[0], stats, M, gamma, beta, 1e-5, 0.1, 0, rm, rv, out=outs[0])
    sums = ops.bn_bwd_reduce(dys[0], outs[0], xs[0], save, relu=True)
    t_apply = timeit(lambda i: ops.bn_apply_train(xs[i % nbuf], stats, M, gamma, beta, 1e-5, 0.1, 0, rm, rv,
                                                  res=rs[i % nbuf] if res else None, out=outs[i % nbuf]), iters)
    k = [0]

    def red(i):
        ops.bn_bwd_reduce(dys[i % nbuf], outs[i % nbuf], xs[i % nbuf], save, relu=True, acc=zs[k[0] * n:(k[0] + 1) * n])
        k[0] += 1
    zs.zero_()
    t_red = timeit(red, iters)
    t_bapply = timeit(lambda i: ops.bn_bwd_apply(dys[i % nbuf], outs[i % nbuf], xs[i % nbuf], save, gamma, sums, M, relu=True, dx=dxs[i % nbuf],
                                                 dres=rs[i % nbuf] if res else None, beta_res=0.0), iters)
    zf = torch.zeros(32 * (iters + 16), device="cuda")
    kf = [0]

    def fused(i, remask=False):
        ops.bn_bwd_fused(dys[i % nbuf], None if remask else outs[i % nbuf], xs[i % nbuf], save, gamma, M, relu=True, dx=dxs[i % nbuf],
                         dres=rs[i % nbuf] if (res and not remask) else None, beta_res=0.0, beta=beta, tickets=zf[kf[0] * 32:kf[0] * 32 + 2])
        kf[0] += 1
    t_fused = timeit(fused, iters)
    zf.zero_()
    kf[0] = 0
    t_fused_rm = timeit(lambda i: fused(i, True), iters)
    b = 2.0 * M * C
    for name, t, passes in (("bn_apply_train", t_apply, 3 if res else 2), ("bn_bwd_reduce", t_red, 3), ("bn_bwd_apply", t_bapply, 5 if res else 4),
                            ("bn_bwd_fused", t_fused, 5 if res else 4), ("bn_bwd_fused(remask,nores)", t_fused_rm, 3)):
        print(f"{name:16s} M={M} C={C} res={res}: {t:7.2f} us  {b * passes / t / 1e6:6.2f} TB/s")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", action="append", required=True)
    ap.add_argument("--res", type=int, default=0)
    ap.add_argument("--iters", type=int, default=60)
    a = ap.parse_args()
    lib.require_device()
    for s in a.shape:
        M, C = (int(v) for v in s.split(","))
        run(M, C, a.res, a.iters)
