"""Worker of tests/test_plugin_dp_gpu.py, launched as `python -m torch.distributed.run --nproc-per-node 2 tests/dp_worker.py OUT GRAPHS`:
what `torchrun -m seg_b200.launch train.py` sets up (rotated device list so that the own GPU is cuda:0, NCCL process group,
`use_synch_bn`), then ONLY the calls an unmodified trainer makes: model(x) -> CrossEntropyLoss2d -> backward -> optimizer.step()."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-segmentation_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    out_path, graphs = sys.argv[1], sys.argv[2] == "1"
    if os.environ.get("SEG_DP_WORKER_DUMP_S"):  # diagnostics: dump every thread's stack if the worker is still alive after N seconds
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["SEG_DP_WORKER_DUMP_S"]), exit=True)
    from seg_b200 import launch
    rank, world = launch.init_data_parallel()  # before any CUDA call of this process
    import torch
    import torch.distributed as dist
    assert torch.cuda.current_device() == 0
    import seg_b200
    from seg_b200.optim import SGD
    from oracle import synth, weights
    sd = weights.deeplab_resnet_state_dict(7, "resnet14", seed=11)
    x, y = synth.make_batch(8, 65, 65, 7, 255, seed=31)
    y[0, 8:40, :] = 255  # unequal valid-pixel counts on the two ranks
    m = seg_b200.DeepLab(7, backbone="resnet14", pretrained=False)
    m.load_state_dict(sd)
    m.engine_dropout = False
    m.use_sync_bn = True  # overlay/utils/sync_batchnorm.convert_model does this for config["use_synch_bn"]
    m = m.cuda().train()
    if graphs:
        m.cuda_graphs(True, warmup=int(os.environ.get("SEG_DP_WORKER_WARMUP", "2")))
    crit = seg_b200.CrossEntropyLoss2d(ignore_index=255)
    opt = SGD([{"params": list(m.get_decoder_params())}, {"params": list(m.get_backbone_params()), "lr": 0.001}], lr=0.01, momentum=0.9, weight_decay=1e-4)
    n = 8 // world
    half = slice(rank * n, rank * n + n)
    xd, yd = x[half].cuda(), y[half].cuda()
    losses = []
    for _ in range(5):
        opt.zero_grad(set_to_none=True)
        loss = crit(m(xd), yd)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert m.bn_sync is not None and m.bn_sync.world == world
    flat = torch.cat([p.detach().reshape(-1) for p in m.parameters()] + [b.detach().float().reshape(-1) for n_, b in m.named_buffers() if "running_" in n_])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    lt = torch.tensor(losses, device="cuda", dtype=torch.float64)
    lg = [torch.empty_like(lt) for _ in range(world)]
    dist.all_gather(lg, lt)
    m.release_graphs()
    torch.cuda.synchronize()
    if rank == 0:
        # single-GPU control on the concatenated batch (no process-group involvement: dp_reduce off, local loss)
        m1 = seg_b200.DeepLab(7, backbone="resnet14", pretrained=False)
        m1.load_state_dict(sd)
        m1.engine_dropout = False
        m1.dp_reduce = False
        m1 = m1.cuda().train()
        from seg_b200.losses import _CEFn
        opt1 = SGD([{"params": list(m1.get_decoder_params())}, {"params": list(m1.get_backbone_params()), "lr": 0.001}], lr=0.01, momentum=0.9, weight_decay=1e-4)
        l1 = []
        for _ in range(5):
            opt1.zero_grad(set_to_none=True)
            loss = _CEFn.apply(m1(x.cuda()), y.cuda(), 255, False)
            loss.backward()
            opt1.step()
            l1.append(loss.item())
        torch.save({"replicas_equal": all(torch.equal(gathered[0], g) for g in gathered[1:]),
                    "losses_equal": all(torch.equal(lg[0], g) for g in lg[1:]), "losses2": losses, "losses1": l1}, out_path)
    dist.barrier()
    torch.cuda.synchronize()
    if graphs:
        # Captured graphs hold NCCL kernels of this communicator; tearing the communicator down while any graph object is
        # still referenced (autograd keeps the last step's entry alive) blocks inside ncclCommDestroy (measured: both ranks
        # parked in destroy_process_group after every check had passed).  A finished worker just exits.
        sys.stdout.flush()
        os._exit(0)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
