"""Graph-replayed plugin path (`model.cuda_graphs()`): `model(x)` / `loss.backward()` replay captured forward / backward
tapes and must be the same training step as the eager tape (`_EngineFn`) from the same state on the same batch.

Every comparison carries an eager-vs-eager control from the same state: split-K wgrad and the BN statistics use fp32
atomics, so two eager runs already differ in the last bits, and a batch-statistics network amplifies that."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import synth, weights

if torch.cuda.is_available():
    import seg_b200
    from seg_b200 import lib


def relerr(a, b):
    a, b = a.detach().double(), b.detach().double()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def cosine(a, b):
    return torch.nn.functional.cosine_similarity(a.detach().double().flatten(), b.detach().double().flatten(), dim=0).item()


def snapshot(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def restore(m, snap):
    with torch.no_grad():
        for k, v in m.state_dict().items():
            v.copy_(snap[k])


def one_step(m, crit, xd, yd, aux_weight=0.4):
    """forward + loss + backward through the plugin surface; returns (logits copy, loss, flat grads, launches)."""
    m.zero_grad(set_to_none=True)
    lib.reset_launch_count()
    out = m(xd)
    n_fwd = lib.launch_count()
    if isinstance(out, tuple):
        loss = crit(out[0], yd) + aux_weight * crit(out[1], yd)
        out = out[0]
    else:
        loss = crit(out, yd)
    out_copy = out.detach().clone()
    lib.reset_launch_count()
    loss.backward()
    n_bwd = lib.launch_count()
    grads = {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}
    return out_copy, float(loss.item()), grads, (n_fwd, n_bwd)


def flat(grads, names):
    return torch.cat([grads[n].reshape(-1) for n in names])


def build(kind, nc=7):
    if kind == "deeplab":
        sd = weights.deeplab_resnet_state_dict(nc, "resnet14", seed=21, randomize_bn=True)
        m = seg_b200.DeepLab(nc, backbone="resnet14", output_stride=16)
        size = 65
    else:
        sd = weights.pspnet_state_dict(nc, "resnet50", seed=21, randomize_bn=True)
        m = seg_b200.PSPNet(nc, backbone="resnet50")
        size = 49
    m.load_state_dict(sd, strict=True)
    m.engine_dropout = False
    return m.cuda().train(), size


def log(gpu_out_dir, msg):
    print(msg)
    with open(f"{gpu_out_dir}/model_parity.txt", "a") as f:
        f.write(msg + "\n")


@pytest.mark.parametrize("kind", ["deeplab", "pspnet"])
def test_graphed_plugin_step_equals_eager_step_frozen_bn(kind, gpu_out_dir):
    """The sharp check.  With BatchNorm frozen (`freeze_bn()`, the reference's `arch.args.freeze_bn`) the forward has no
    atomics and the network is well conditioned, so the replayed tapes must reproduce the eager tape's logits exactly and
    its gradients up to the fp32-atomic order of split-K wgrad.  PSPNet also exercises the (out, aux) pair."""
    nc = 7
    m, size = build(kind, nc)
    m.freeze_bn()
    crit = seg_b200.CrossEntropyLoss2d(ignore_index=255)
    # frozen random running statistics make the loss and its gradients huge, so the step between iterations is
    # normalised: every parameter moves by 0.1 % of its own norm along its gradient — finite, and large enough that a
    # replay packing stale weights would show in the logits
    opt = torch.optim.SGD(m.parameters(), lr=1.0)
    m.cuda_graphs(True, warmup=1)
    m._graphs_enabled = False  # toggled per run below; cuda_graphs(False) would also release the captured graphs
    replayed = 0
    prev_out = None
    for i in range(4):
        x, y = synth.make_batch(2, size, size, nc, 255, seed=9100)  # same batch every step: logits change only through the weights
        xd, yd = x.cuda(), y.cuda()
        snap = snapshot(m)
        out_a, loss_a, g_a, _ = one_step(m, crit, xd, yd)
        restore(m, snap)
        out_b, loss_b, g_b, _ = one_step(m, crit, xd, yd)          # control: eager again from the same state
        restore(m, snap)
        m._graphs_enabled = True
        out_g, loss_g, g_g, (n_fwd, n_bwd) = one_step(m, crit, xd, yd)  # i = 0: eager warm-up, 1: capture + replay, 2..: replay
        m._graphs_enabled = False
        names = sorted(g_a)
        assert sorted(g_g) == names == sorted(g_b), "graph path returned gradients for a different parameter set"
        noise_o = relerr(out_b, out_a)
        c_ctrl, c_graph = cosine(flat(g_b, names), flat(g_a, names)), cosine(flat(g_g, names), flat(g_a, names))
        log(gpu_out_dir, f"[plugin-graph frozen-BN {kind}] step {i}: logits relerr {relerr(out_g, out_a):.2e} (control {noise_o:.2e}) "
            f"loss {loss_g:.6f} vs {loss_a:.6f} grad cosine {c_graph:.6f} (control {c_ctrl:.6f}) launches fwd/bwd {n_fwd}/{n_bwd}")
        assert torch.isfinite(out_a).all() and loss_a == loss_a
        assert relerr(out_g, out_a) <= max(1e-6, 2 * noise_o)
        assert abs(loss_g - loss_a) <= 1e-5 * abs(loss_a)
        assert c_graph >= min(0.9999, c_ctrl - 1e-3)
        if prev_out is not None:  # the optimiser step between iterations must be visible to the replayed forward
            assert relerr(out_g, prev_out) > 1e-5, "logits did not move after optimizer.step(): stale packed weights?"
        prev_out = out_g
        if i >= 2:  # pure replay: the model's kernels are launched by cudaGraphLaunch, not through the C ABI
            assert n_fwd == 0 and n_bwd <= 8, (n_fwd, n_bwd)  # backward: only the CE kernels of the eager loss
            replayed += 1
        with torch.no_grad():
            for p in m.parameters():
                if p.grad is not None:
                    p.grad.mul_(1e-3 * p.norm() / (p.grad.norm() + 1e-20))
        opt.step()  # from the graph run's gradients: the next iteration starts from new weights
    assert replayed == 2 and len(m._graph_entries) == 1


def test_graphed_plugin_step_batch_statistics(gpu_out_dir):
    """Batch-statistics BatchNorm through the replayed tapes: running statistics, num_batches_tracked and the loss follow
    the eager tape.  A random-init network with batch-stat BN on 5x5 maps amplifies the fp32-atomic order of the
    statistics sums (eager vs eager from the same state differs by up to several percent in the logits — the control), so
    the bounds scale with the control."""
    nc = 7
    m, size = build("deeplab", nc)
    crit = seg_b200.CrossEntropyLoss2d(ignore_index=255)
    m.cuda_graphs(True, warmup=1)
    m._graphs_enabled = False
    for i in range(4):
        x, y = synth.make_batch(2, size, size, nc, 255, seed=9150 + i)
        xd, yd = x.cuda(), y.cuda()
        snap = snapshot(m)
        out_a, loss_a, g_a, _ = one_step(m, crit, xd, yd)
        stats_a = snapshot(m)
        restore(m, snap)
        out_b, loss_b, g_b, _ = one_step(m, crit, xd, yd)
        stats_b = snapshot(m)
        restore(m, snap)
        m._graphs_enabled = True
        out_g, loss_g, g_g, _ = one_step(m, crit, xd, yd)
        m._graphs_enabled = False
        stats_g = snapshot(m)
        noise_o, noise_l = relerr(out_b, out_a), abs(loss_b - loss_a) / abs(loss_a)
        worst_ctrl = worst = 0.0
        for k in stats_a:
            if k.endswith("running_mean") or k.endswith("running_var"):
                worst_ctrl = max(worst_ctrl, relerr(stats_b[k], stats_a[k]))
                worst = max(worst, relerr(stats_g[k], stats_a[k]))
            if k.endswith("num_batches_tracked"):
                assert int(stats_g[k]) == int(stats_a[k]) == int(snap[k]) + 1, k
        log(gpu_out_dir, f"[plugin-graph batch-stat deeplab] step {i}: logits relerr {relerr(out_g, out_a):.2e} (control {noise_o:.2e}) "
            f"loss {loss_g:.6f} vs {loss_a:.6f} (control {loss_b:.6f}) running stats {worst:.2e} (control {worst_ctrl:.2e})")
        assert all(torch.isfinite(v).all() for v in g_g.values()) and sorted(g_g) == sorted(g_a)
        # sanity bounds only (the frozen-BN test above is the sharp one): at batch 2 the ASPP image-pooling branch
        # normalises TWO samples per channel, a sign function that turns last-bit differences into O(1) changes
        assert relerr(out_g, out_a) <= max(0.15, 6 * noise_o)
        assert abs(loss_g - loss_a) <= max(2e-2, 6 * noise_l) * abs(loss_a)
        assert worst <= max(5e-2, 6 * worst_ctrl)


def test_graphed_eval_forward_and_release():
    nc = 5
    sd = weights.deeplab_resnet_state_dict(nc, "resnet14", seed=4, randomize_bn=True)
    m = seg_b200.DeepLab(nc, backbone="resnet14", output_stride=16)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    x, _ = synth.make_batch(2, 65, 65, nc, 255, seed=9200)
    xd = x.cuda()
    with torch.no_grad():
        ref = m(xd).clone()
        m.cuda_graphs(True, warmup=1)
        outs = [m(xd).clone() for _ in range(3)]  # eager, capture + replay, replay
        x2 = torch.flip(xd, dims=[3])
        flipped = m(x2).clone()                   # same shape -> replay with new contents
        m.cuda_graphs(False)
        flipped_ref = m(x2)
    for o in outs:
        assert relerr(o, ref) < 1e-6, "eval forward is deterministic: the replay must reproduce the eager logits"
    assert relerr(flipped, flipped_ref) < 1e-6
    assert relerr(flipped, ref) > 1e-3, "the replay did not see the new input"
    assert not m._graph_entries
