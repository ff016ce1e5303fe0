"""The plugin surface as a multi-GPU drop-in: see tests/dp_worker.py (one process per GPU under torchrun,
exactly as `torchrun --nproc-per-node 2 -m seg_b200.launch train.py` would set things up).  Needs 2 GPUs."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("graphs", [False, True], ids=["eager", "graph-replay"])
def test_plugin_surface_two_gpus_replicas_stay_identical(tmp_path, graphs):
    out = str(tmp_path / "r.pt")
    port = 29900 + (os.getpid() + int(graphs) * 13) % 1000
    env = {k: v for k, v in os.environ.items() if k not in ("CUDA_VISIBLE_DEVICES", "SEG_DEVICES_ROTATED")}
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(HERE, "dp_worker.py"), out, "1" if graphs else "0"],
                       env=env, capture_output=True, text=True, timeout=280)
    assert r.returncode == 0 and os.path.exists(out), r.stdout[-2000:] + r.stderr[-4000:]
    res = torch.load(out)
    print(res)
    assert res["replicas_equal"], "replicas diverged (parameters / running statistics differ between the two ranks)"
    assert res["losses_equal"], "the ranks report different losses: the loss is not the global-batch mean"
    # same training as one GPU on the concatenated batch (SyncBN + global mean + gradient mean), up to bf16 / clamp(var, eps) noise
    assert abs(res["losses2"][0] - res["losses1"][0]) < 2e-2 * abs(res["losses1"][0]), res
    assert res["losses2"][-1] < res["losses2"][0] and abs(res["losses2"][-1] - res["losses1"][-1]) < 0.1 * abs(res["losses1"][-1]), res
