"""CPU restatement of the reference's per-step segmentation metrics — TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's
cpu_baseline leg); the product never imports it.  Pinned by tests/golden/metrics.npz (generated from the reference by
oracle/make_golden_metrics.py).

Follows utils/metrics.py:
  eval_metrics               :59-67  predict = argmax+1, target+1, labeled = (target>0)&(target<=num_class)
  batch_pix_accuracy         :41-45  correct = sum((predict==target)*labeled), labeled = sum(labeled)
  batch_intersection_union   :47-57  predict*=labeled; inter = predict*(predict==target); three histc(bins=K, min=1, max=K);
                                     union = pred + lab - inter
torch.histc with bins=K over [1, K] puts the integer value v in bin floor((v-1)*K/(K-1)) = v-1 (last bin closed), i.e. plain
per-class counts; values outside [1, K] (0 = unlabeled, ignore_index+1) are dropped.
"""
import numpy as np


def eval_metrics(output, target, num_class):
    """output: float array [N, C, H, W]; target: int array [N, H, W].  Returns [correct, labeled, inter[K], union[K]]."""
    output = np.asarray(output)
    target = np.asarray(target).astype(np.int64)
    predict = output.argmax(1).astype(np.int64) + 1  # first maximum, like torch.max
    tgt = target + 1
    labeled = (tgt > 0) & (tgt <= num_class)
    correct = int(((predict == tgt) & labeled).sum())
    n_labeled = int(labeled.sum())
    predict = predict * labeled
    inter_map = predict * (predict == tgt)

    def hist(v):
        v = v.reshape(-1)
        v = v[(v >= 1) & (v <= num_class)]
        return np.bincount(v - 1, minlength=num_class).astype(np.float32)

    inter, pred, lab = hist(inter_map), hist(predict), hist(tgt)
    union = pred + lab - inter
    return [np.round(np.asarray(correct), 5), np.round(np.asarray(n_labeled), 5), np.round(inter, 5), np.round(union, 5)]
