"""CPU restatement of the tail of the reference's sample pipeline — base/base_dataset.py:93-123 (zero-pad to the crop
size, crop, horizontal flip) and :129-136 (label -> int64, ToTensor, Normalize).  TEST INFRASTRUCTURE ONLY.
Pinned by tests/golden/data_tail.npz (oracle/make_golden_data.py runs the reference's own BaseDataSet.__getitem__)."""
import numpy as np
import torch


def sample_tail(image, label, crop_size, y0, x0, flip, mean, std):
    """image uint8 [h,w,3], label [h,w] -> (fp32 [3,crop,crop], int64 [crop,crop])"""
    h, w, _ = image.shape
    pad_h, pad_w = max(crop_size - h, 0), max(crop_size - w, 0)
    if pad_h > 0 or pad_w > 0:  # cv2.copyMakeBorder(top=0, bottom=pad_h, left=0, right=pad_w, BORDER_CONSTANT, value=0)
        image = np.pad(image, ((0, pad_h), (0, pad_w), (0, 0)), mode="constant", constant_values=0)
        label = np.pad(label, ((0, pad_h), (0, pad_w)), mode="constant", constant_values=0)
    image = image[y0:y0 + crop_size, x0:x0 + crop_size]
    label = label[y0:y0 + crop_size, x0:x0 + crop_size]
    if flip:
        image = np.fliplr(image).copy()
        label = np.fliplr(label).copy()
    lab = torch.from_numpy(np.array(label, dtype=np.int32)).long()
    t = torch.from_numpy(np.ascontiguousarray(np.uint8(image))).permute(2, 0, 1).contiguous().to(torch.float32).div(255)  # ToTensor
    m = torch.as_tensor(mean, dtype=torch.float32).view(-1, 1, 1)
    s = torch.as_tensor(std, dtype=torch.float32).view(-1, 1, 1)
    return t.sub_(m).div_(s), lab  # Normalize
