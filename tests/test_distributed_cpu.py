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
    m = nets.DeepLab(7, backbone="resnet14", output_stride=16)
    m.load_state_dict(sd, strict=True)
    m.engine_dropout = False
    m.bn_sync = sync
    m.train()
    out = m(x)
    loss = plosses._CEFn.apply(out, y, 255)
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
