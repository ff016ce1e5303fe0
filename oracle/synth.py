"""Synthetic inputs of SURVEY.md §8(d): randn images; 16x16-block-constant labels with a 4-px ignore border."""
import torch


def make_batch(B, H, W, num_classes, ignore_index=255, seed=1234):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 3, H, W, generator=g)
    bh, bw = (H + 15) // 16, (W + 15) // 16
    blocks = torch.randint(0, num_classes, (B, bh, bw), generator=g)
    y = blocks.repeat_interleave(16, 1).repeat_interleave(16, 2)[:, :H, :W].contiguous()
    y[:, :4, :] = ignore_index
    y[:, -4:, :] = ignore_index
    y[:, :, :4] = ignore_index
    y[:, :, -4:] = ignore_index
    return x, y
