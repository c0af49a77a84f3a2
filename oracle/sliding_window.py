"""CPU restatement of MONAI 0.7.0 ``monai.inferers.sliding_window_inference`` (constant blend).

TEST INFRASTRUCTURE ONLY (see oracle/ltu_oracle.py for the rules).

PARITY UNPINNED: the arithmetic of this step lives in a third-party dependency of the reference
(MONAI, pinned to 0.7.0 in requirements.txt:1; call sites inference_multi_classes.py:143,
inference_embed_attn.py:141, utils/utils_3D_embed_full.py:148, utils/utils_3D_multi_class.py:184).
MONAI is neither vendored under /root/reference nor installed/installable in this image, and the
reference holds no test or golden vector for it, so this file restates the published algorithm
(inferers/utils.py::sliding_window_inference, ::_get_scan_interval, data/utils.py::
dense_patch_slices, ::compute_importance_map for mode="constant") and is anchored only on the
reference's call sites (roi, sw_batch_size, overlap, sigma_scale=0 -> constant mode).
"""
from __future__ import annotations

import math
from typing import Callable, Sequence

import torch
import torch.nn.functional as F


def get_scan_interval(image_size, roi_size, overlap):
    out = []
    for i in range(len(image_size)):
        if roi_size[i] == image_size[i]:
            out.append(int(roi_size[i]))
        else:
            interval = int(roi_size[i] * (1 - overlap))
            out.append(interval if interval > 0 else 1)
    return tuple(out)


def dense_patch_starts(image_size, patch_size, scan_interval):
    nd = len(image_size)
    scan_num = []
    for i in range(nd):
        if scan_interval[i] == 0:
            scan_num.append(1)
        else:
            num = int(math.ceil(float(image_size[i]) / scan_interval[i]))
            scan_dim = next((d for d in range(num) if d * scan_interval[i] + patch_size[i] >= image_size[i]), None)
            scan_num.append(scan_dim + 1 if scan_dim is not None else 1)
    starts = []
    for dim in range(nd):
        dim_starts = []
        for idx in range(scan_num[dim]):
            s = idx * scan_interval[dim]
            s -= max(s + patch_size[dim] - image_size[dim], 0)
            dim_starts.append(s)
        starts.append(dim_starts)
    grid = torch.cartesian_prod(*[torch.tensor(s) for s in starts]).reshape(-1, nd)     # C order ("ij")
    return [tuple(int(v) for v in row) for row in grid]


def sliding_window_inference(inputs: torch.Tensor, roi_size: Sequence[int], sw_batch_size: int,
                             predictor: Callable[[torch.Tensor], torch.Tensor], overlap: float = 0.25,
                             cval: float = 0.0) -> torch.Tensor:
    """inputs [B, Cin, H, W, D] -> fp32 [B, Cout, H, W, D]; constant importance map (all ones)."""
    nd = inputs.dim() - 2
    image_size_ = list(inputs.shape[2:])
    batch = inputs.shape[0]
    roi = tuple(int(r) if r and r > 0 else image_size_[i] for i, r in enumerate(roi_size))
    image_size = tuple(max(image_size_[i], roi[i]) for i in range(nd))
    pad_size = []
    for k in range(inputs.dim() - 1, 1, -1):
        diff = max(roi[k - 2] - inputs.shape[k], 0)
        half = diff // 2
        pad_size.extend([half, diff - half])
    inputs = F.pad(inputs, pad=pad_size, mode="constant", value=cval)
    starts = dense_patch_starts(image_size, roi, get_scan_interval(image_size, roi, overlap))
    num_win = len(starts)
    total = num_win * batch
    out = cnt = None
    for g0 in range(0, total, sw_batch_size):
        idxs = range(g0, min(g0 + sw_batch_size, total))
        slices = []
        for idx in idxs:
            b, s = idx // num_win, starts[idx % num_win]
            slices.append((slice(b, b + 1), slice(None)) + tuple(slice(s[d], s[d] + roi[d]) for d in range(nd)))
        data = torch.cat([inputs[sl] for sl in slices])
        pred = predictor(data).to(torch.float32)
        if out is None:
            shape = [batch, pred.shape[1]] + list(image_size)
            out = torch.zeros(shape, dtype=torch.float32)
            cnt = torch.zeros(shape, dtype=torch.float32)
        for j, sl in enumerate(slices):
            out[sl] += pred[j]
            cnt[sl] += 1.0
    out = out / cnt
    final = [slice(None), slice(None)]
    for sp in range(nd):
        s0 = pad_size[(nd - 1 - sp) * 2]
        final.append(slice(s0, s0 + image_size_[sp]))
    return out[tuple(final)]
