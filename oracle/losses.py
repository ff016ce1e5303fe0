"""CPU restatement of the reference's per-pixel losses (TEST INFRASTRUCTURE).

  cross_entropy2d -> utils/losses.py:24-31 (nn.CrossEntropyLoss(ignore_index, reduction='mean'))
  dice_loss       -> utils/losses.py:33-50 (mutates `target` in place exactly like the reference, :40-42)
  ce_dice_loss    -> utils/losses.py:67-77 (DiceLoss() built with its default ignore_index=255, :71)
  lovasz_softmax  -> utils/losses.py:79-89 + utils/lovasz_losses.py:153-218,19-31 (classes='present', per_image=False)
"""
import torch
import torch.nn.functional as F


def cross_entropy2d(output, target, ignore_index=255):
    return F.cross_entropy(output, target, ignore_index=ignore_index, reduction="mean")


def dice_loss(output, target, smooth=1.0, ignore_index=255):
    # losses.py:40-42 — `ignore_index not in range(min, max)` then overwrite ignored labels with target.min()
    if ignore_index not in range(int(target.min()), int(target.max())):
        if (target == ignore_index).sum() > 0:
            target[target == ignore_index] = target.min()
    one_hot = torch.zeros(output.shape, dtype=torch.float32).scatter_(1, target.unsqueeze(1), 1)
    prob = F.softmax(output, dim=1)
    inter = (prob.reshape(-1) * one_hot.reshape(-1)).sum()
    return 1 - ((2.0 * inter + smooth) / (prob.sum() + one_hot.sum() + smooth))


def ce_dice_loss(output, target, ignore_index=255):
    ce = F.cross_entropy(output, target, ignore_index=ignore_index, reduction="mean")
    return ce + dice_loss(output, target)  # DiceLoss() default ignore 255 (losses.py:71)


def _lovasz_grad(gt_sorted):
    """lovasz_losses.py:19-31."""
    p = gt_sorted.numel()
    gts = gt_sorted.sum()
    inter = gts - gt_sorted.float().cumsum(0)
    union = gts + (1 - gt_sorted).float().cumsum(0)
    jac = 1.0 - inter / union
    if p > 1:
        jac[1:p] = jac[1:p] - jac[0:-1]
    return jac


def lovasz_softmax(output, target, ignore_index=255):
    """LovaszSoftmax.forward (losses.py:86-89) -> lovasz_softmax(per_image=False, classes='present')."""
    prob = F.softmax(output, dim=1)
    B, C, H, W = prob.shape
    prob = prob.permute(0, 2, 3, 1).reshape(-1, C)
    lab = target.reshape(-1)
    valid = lab != ignore_index
    prob, lab = prob[valid], lab[valid]
    if prob.numel() == 0:
        return prob * 0.0
    losses = []
    for c in range(C):
        fg = (lab == c).float()
        if fg.sum() == 0:
            continue
        err = (fg - prob[:, c]).abs()
        err_sorted, perm = torch.sort(err, 0, descending=True)
        losses.append(torch.dot(err_sorted, _lovasz_grad(fg[perm])))
    return sum(losses) / len(losses)


def syncbn_mean_istd(sum_, ssum, size, eps=1e-5):
    """sync_batchnorm/batchnorm.py:128-145 (_compute_mean_std): inv_std = clamp(var, eps)^-1/2, unbiased var for
    the running estimate.  Returns (mean, inv_std, unbiased_var)."""
    mean = sum_ / size
    sumvar = ssum - sum_ * mean
    unbias_var = sumvar / (size - 1)
    bias_var = sumvar / size
    return mean, bias_var.clamp(eps) ** -0.5, unbias_var
