"""N>1 host logic on CPU (gloo, world_size 2): with SyncBN on, a 2-rank step on half batches must equal the 1-rank
step on the concatenated batch — the property the reference's SyncBN is built for (sync_batchnorm/batchnorm.py:160-167)
— for the loss, the rank-averaged parameter gradients and the BN running statistics.  Kernels are the ATen emulation
(fp32 storage); the statistics exchange goes through torch.distributed instead of the NVLink kernel (same interface:
`allreduce_(vec)` sums an fp32 vector over ranks in place)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


class GlooSync:
    def __init__(self):
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def allreduce_(self, vec):
        dist.all_reduce(vec)
        return vec


def _setup_emulation():
    for p in (ROOT, os.path.join(ROOT, "pytorch-segmentation_b200"), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import cpu_emulation as emu
    from seg_b200 import engine, nets
    from seg_b200 import losses as plosses
    for mod in (engine, nets, plosses):
        mod.ops = emu
    engine.ACT_DTYPE = torch.float32
    emu.ACT_DTYPE = torch.float32
    nets._EngineModel._check_input = lambda self, x: None
    return nets, plosses


def _step(nets, plosses, sd, x, y, sync):
    """One forward/backward with the engine's own data-parallel exchange OFF (dp_reduce=False, per-rank mean loss): this
    test does the gradient averaging by hand, and its single-process control runs while the group is still alive."""
    m = nets.DeepLab(7, backbone="resnet14", output_stride=16)
    m.load_state_dict(sd, strict=True)
    m.engine_dropout = False
    m.bn_sync = sync
    m.dp_reduce = False
    m.train()
    out = m(x)
    loss = plosses._CEFn.apply(out, y, 255, False)
    loss.backward()
    return m, loss.detach()


def _worker(rank, world, port, result_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    nets, plosses = _setup_emulation()
    from oracle import synth, weights
    sd = weights.deeplab_resnet_state_dict(7, "resnet14", seed=11, randomize_bn=True)
    x, y = synth.make_batch(4, 49, 49, 7, 255, seed=31)
    half = slice(rank * 2, rank * 2 + 2)
    m, loss = _step(nets, plosses, sd, x[half].contiguous(), y[half].contiguous(), GlooSync())
    grads = torch.cat([p.grad.reshape(-1) for p in m.parameters()])
    dist.all_reduce(grads)
    grads /= world
    dist.all_reduce(loss)
    loss /= world
    stats = torch.cat([b.reshape(-1).float() for n, b in m.named_buffers() if "running_" in n])
    if rank == 0:
        m1, loss1 = _step(nets, plosses, sd, x, y, None)  # single process, concatenated batch
        g1 = torch.cat([p.grad.reshape(-1) for p in m1.parameters()])
        s1 = torch.cat([b.reshape(-1).float() for n, b in m1.named_buffers() if "running_" in n])
        torch.save({"loss2": loss, "loss1": loss1,
                    "cos": torch.nn.functional.cosine_similarity(grads.double(), g1.double(), dim=0),
                    "grad_rel": (grads - g1).abs().max() / g1.abs().max(),
                    "stats_rel": (stats - s1).abs().max() / s1.abs().max()}, result_path)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_syncbn_step_equals_single_rank_on_concatenated_batch(tmp_path):
    result = str(tmp_path / "r.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, result), nprocs=2, join=True)
    r = torch.load(result)
    assert abs(r["loss2"].item() - r["loss1"].item()) < 1e-4 * abs(r["loss1"].item()), r
    assert r["stats_rel"].item() < 1e-4, r
    assert r["cos"].item() > 0.999 and r["grad_rel"].item() < 5e-2, r


def _plugin_worker(rank, world, port, result_path):
    """The plugin surface under a process group: model(x) -> CrossEntropyLoss2d -> backward -> torch.optim.SGD, nothing
    else — the gradient exchange and the global-mean loss live inside the engine's autograd nodes (what an unmodified
    train.py gets from `torchrun -m seg_b200.launch`)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    nets, plosses = _setup_emulation()
    from oracle import synth, weights
    sd = weights.deeplab_resnet_state_dict(7, "resnet14", seed=11, randomize_bn=True)
    x, y = synth.make_batch(4, 49, 49, 7, 255, seed=31)
    y[0, 8:30, :] = 255  # unequal valid-pixel counts per rank: the global mean differs from the mean of per-rank means
    half = slice(rank * 2, rank * 2 + 2)

    def train(xs, ys, sync, steps=3):
        m = nets.DeepLab(7, backbone="resnet14", output_stride=16)
        m.load_state_dict(sd, strict=True)
        m.engine_dropout = False
        m.bn_sync = sync
        m.train()
        opt = torch.optim.SGD([{"params": list(m.get_decoder_params())}, {"params": list(m.get_backbone_params()), "lr": 0.001}],
                              lr=0.01, momentum=0.9, weight_decay=1e-4)
        crit = plosses.CrossEntropyLoss2d(ignore_index=255)
        crit.forward = lambda o, t: plosses._CEFn.apply(o, t, 255, True)  # the CUDA check of the module is not for this emulation
        losses = []
        for _ in range(steps):
            opt.zero_grad()
            loss = crit.forward(m(xs), ys)
            loss.backward()
            opt.step()
            losses.append(loss.item())
        return m, losses

    m, losses = train(x[half].contiguous(), y[half].contiguous(), GlooSync())
    flat = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        dist.destroy_process_group()  # the single-process control below must see world == 1
        m1, losses1 = train(x, y, None)
        flat1 = torch.cat([p.detach().reshape(-1) for p in m1.parameters()])
        torch.save({"replicas_equal": all(torch.equal(gathered[0], g) for g in gathered[1:]), "losses2": losses, "losses1": losses1,
                    "param_rel": ((flat - flat1).abs().max() / flat1.abs().max()).item(),
                    "upd_cos": torch.nn.functional.cosine_similarity((flat - torch.cat([sd[n].reshape(-1) for n, _ in m.named_parameters()])).double(),
                                                                     (flat1 - torch.cat([sd[n].reshape(-1) for n, _ in m1.named_parameters()])).double(), dim=0).item()},
                   result_path)
    else:
        dist.destroy_process_group()


def test_plugin_surface_under_process_group_keeps_replicas_identical(tmp_path):
    """with one process per GPU the plugin surface itself must exchange gradients (nn.DataParallel did it
    in the reference, base/base_trainer.py:33-38): after 3 SGD steps the replicas are BIT-equal, the reported loss is the
    global mean over both ranks' valid pixels, and the trajectory equals the single-process run on the concatenated batch."""
    result = str(tmp_path / "r.pt")
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_plugin_worker, args=(2, port, result), nprocs=2, join=True)
    r = torch.load(result)
    assert r["replicas_equal"], "data-parallel replicas diverged"
    for a, b in zip(r["losses2"], r["losses1"]):
        assert abs(a - b) < 2e-3 * abs(b), r  # fp32 summation order (2 x half batch vs one batch), amplified by 3 SGD steps
    assert r["upd_cos"] > 0.999 and r["param_rel"] < 1e-3, r
