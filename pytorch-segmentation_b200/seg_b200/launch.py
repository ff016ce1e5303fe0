"""Run an UNMODIFIED reference script (train.py / inference.py) with the B200-native models and losses plugged in:

    cd /path/to/pytorch-segmentation
    PYTHONPATH=/path/to/repo/pytorch-segmentation_b200 python -m seg_b200.launch train.py -c config.json

`python train.py` puts the script's own directory first on sys.path, ahead of PYTHONPATH, so the overlay packages
(`overlay/models`, `overlay/utils`) could never win there; this launcher orders sys.path as
[overlay, reference root, ...] and then executes the script as __main__ — the script's source is untouched.
Under torchrun (one process per GPU) it also restricts each process to its own device so the reference's
`n_gpu: 1` DataParallel-of-one path (base/base_trainer.py:33-38) runs one replica per GPU.
"""
import os
import runpy
import sys


def setup_paths(reference_root):
    here = os.path.dirname(os.path.abspath(__file__))
    pkg = os.path.dirname(here)
    overlay = os.path.join(pkg, "overlay")
    reference_root = os.path.abspath(reference_root)
    os.environ.setdefault("SEG_REFERENCE_ROOT", reference_root)
    for p in (reference_root, pkg, overlay):  # final order: overlay, pkg, reference root
        while p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    # anything of the reference imported before the path was ordered must be re-resolved through the overlay
    for name in list(sys.modules):
        if name.split(".")[0] in ("models", "utils", "base"):
            del sys.modules[name]
    # utils/transforms.py:4 imports skimage, which this image lacks; the trainer never calls it (SURVEY.md §8c shim 1)
    try:
        import skimage  # noqa: F401
    except Exception:
        import types
        for n in ("skimage", "skimage.filters"):
            sys.modules.setdefault(n, types.ModuleType(n))
        sys.modules["skimage.filters"].gaussian = None


def main():
    if len(sys.argv) < 2:
        print(__doc__)
        sys.exit(2)
    script = os.path.abspath(sys.argv[1])
    setup_paths(os.path.dirname(script))
    if "LOCAL_RANK" in os.environ and "CUDA_VISIBLE_DEVICES" not in os.environ:
        os.environ["CUDA_VISIBLE_DEVICES"] = os.environ["LOCAL_RANK"]
    if os.environ.get("SEG_FUSED_OPTIM", "1") != "0":
        # base/base_trainer.py:57 builds getattr(torch.optim, config['optimizer']['type']): "SGD" over CUDA fp32
        # parameters resolves to the torch.optim.SGD subclass whose step() is the multi-tensor kernel (same param_groups /
        # state_dict / schedulers); any other use gets the stock class
        from seg_b200 import optim
        optim.install()
    sys.argv = [script] + sys.argv[2:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
