"""CPU restatement of what the reference's inference scripts do to the stitched volume
(inference_embed_attn.py:141-160, inference_multi_classes.py:143-162) and of the evaluation
metrics they print.

TEST INFRASTRUCTURE ONLY (see oracle/ltu_oracle.py for the rules): imported by tests/ only.

Pinning:
* ``decide_*``, the metric functions (``dice_class_loss`` ... ``localization_loss_multi``):
  PINNED -- tools/make_golden_postproc.py imports the unmodified classes of loss/criterions.py and
  loss/multi_criterions.py from /root/reference and stores their values on seeded inputs in
  tests/golden/postproc.npz (tests/test_postproc_cpu.py).
* ``keep_largest_connected_component``: PARITY UNPINNED -- the arithmetic lives in MONAI 0.7.0
  (requirements.txt:1; monai/transforms/post/array.py::KeepLargestConnectedComponent and
  monai/transforms/utils.py::get_largest_connected_component_mask, which calls
  skimage.measure.label), neither vendored nor installable here and covered by no reference test.
  Restated from the published algorithm with scipy.ndimage.label (same raster-order component
  numbering as skimage.measure.label); anchored on the call site inference_multi_classes.py:104
  (applied_labels=[1, 2], independent=False, connectivity=3) and :150.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch


# ----------------------------------------------------------------------------- stitched volume
def vote_fractions(votes: np.ndarray) -> np.ndarray:
    """uint8 votes [C,...] -> fp32 fractions: MONAI's ``output_image / count_map`` for one-hot windows."""
    v = votes.astype(np.float32)
    return v / v.sum(0, keepdims=True, dtype=np.float32)


def decide_threshold(frac: np.ndarray, thr: float = 0.5) -> np.ndarray:
    """inference_embed_attn.py:147 ``(predict >= threshold).float()``."""
    return (frac >= np.float32(thr)).astype(np.uint8)


def decide_round(frac: np.ndarray) -> np.ndarray:
    """inference_multi_classes.py:148 ``torch.round(predict)`` (half to even)."""
    return torch.round(torch.from_numpy(np.ascontiguousarray(frac))).numpy().astype(np.uint8)


def largest_component_mask(fg: np.ndarray, connectivity: int) -> np.ndarray:
    """monai.transforms.utils.get_largest_connected_component_mask for one item (0.7.0):
    ``labels = measure.label(fg, connectivity)``; ``labels == argmax(bincount(labels.flat)[1:]) + 1``."""
    from scipy import ndimage
    structure = ndimage.generate_binary_structure(fg.ndim, connectivity)
    labels, n = ndimage.label(fg != 0, structure=structure)
    if n == 0:
        return np.zeros_like(fg, dtype=bool)
    return labels == (np.argmax(np.bincount(labels.ravel())[1:]) + 1)


def keep_largest_connected_component(onehot: np.ndarray, applied_labels: Sequence[int], independent: bool = False,
                                     connectivity: int = 3) -> np.ndarray:
    """KeepLargestConnectedComponent.__call__, one-hot branch (img.shape[0] > 1), MONAI 0.7.0."""
    out = onehot.copy()
    if independent:
        for i in applied_labels:
            fg = out[i] != 0
            mask = largest_component_mask(fg, connectivity)
            out[i][fg != mask] = 0
        return out
    fg = np.any(out[list(applied_labels)] != 0, axis=0)
    mask = largest_component_mask(fg, connectivity)
    drop = fg != mask
    for i in applied_labels:
        out[i][drop] = 0
    return out


def background_from_rest(onehot: np.ndarray) -> np.ndarray:
    """inference_multi_classes.py:152 ``predict2[:, 0] = 1 - predict2[:, 1] - predict2[:, 2]`` (general C)."""
    out = onehot.astype(np.int16)
    out[0] = 1 - out[1:].sum(0)
    return out


# ----------------------------------------------------------------------------- metrics (binary scripts)
def _class_vectors(predict: torch.Tensor, target: torch.Tensor, class_index: int):
    """loss/criterions.py:52-61: predict [N,C,...] float, target [N,1,...] (0/1 labels used as the class-1 mask)."""
    p = predict.flatten(2).transpose(2, 1)[:, :, class_index]
    t = target.flatten(2).transpose(2, 1).squeeze(2)
    return p, t


def dice_class_loss(predict, target, class_index=1, eps=1e-9):
    """loss/criterions.py:46-69."""
    p, t = _class_vectors(predict, target, class_index)
    return 1 - torch.mean((2 * torch.sum(p * t, -1) + eps) / (torch.sum(p + t, -1) + eps))


def recall(predict, target, class_index=1, eps=1e-5):
    """loss/criterions.py:291-311."""
    p, t = _class_vectors(predict, target, class_index)
    return torch.mean((torch.sum(p * t, -1) + eps) / (torch.sum(t, -1) + eps))


def precision(predict, target, class_index=1, eps=1e-5):
    """loss/criterions.py:359-379."""
    p, t = _class_vectors(predict, target, class_index)
    return torch.mean((torch.sum(p * t, -1) + eps) / (torch.sum(p, -1) + eps))


def _profile_distance(p_prof: torch.Tensor, t_prof: torch.Tensor, eps: float, scale: float):
    """dis_loss (loss/criterions.py:231-241 with the factor 8, loss/multi_criterions.py:271-281 without)."""
    dp = torch.cumsum(p_prof, -1) / (torch.sum(p_prof, -1, keepdim=True) + eps)
    dt = torch.cumsum(t_prof, -1) / (torch.sum(t_prof, -1, keepdim=True) + eps)
    return scale * torch.mean(torch.abs(dp - dt))


def localization_loss(predict, target, class_index=1, eps=1e-6, mask_threshold=10):
    """loss/criterions.py:192-228.  All three loop iterations reduce over (W, D): `transpose(2, 2)` is the identity for
    i = 0 and i = 1, 2 flatten the same axes, so the value is the H-profile distance (three equal terms / 3)."""
    p = predict[:, class_index].unsqueeze(1).float()
    t = target.float()
    pp = torch.sigmoid(p.flatten(3).sum(-1) - mask_threshold)
    tp = torch.sigmoid(t.flatten(3).sum(-1) - mask_threshold)
    return _profile_distance(pp, tp, eps, 8.0)


# ----------------------------------------------------------------------------- metrics (multi-class script)
def dice_class_loss_multi(predict, target_onehot, class_index, eps=1e-9):
    """loss/multi_criterions.py:69-83 (class_index 1), :96-110 (2); class_index 0 = DiceClassLoss0 :41-55, which
    scores the FOREGROUND 1 - channel 0."""
    p = predict.flatten(2).transpose(2, 1)
    t = target_onehot.flatten(2).transpose(2, 1)
    if class_index == 0:
        cp, ct = 1 - p[:, :, 0], 1 - t[:, :, 0]
    else:
        cp, ct = p[:, :, class_index], t[:, :, class_index]
    return 1 - torch.mean((2 * torch.sum(cp * ct, -1) + eps) / (torch.sum(cp + ct, -1) + eps))


def recall_multi(predict, target_onehot, class_index, eps=1e-5):
    """loss/multi_criterions.py Recall / Recall2 (:359-374)."""
    p = predict.flatten(2).transpose(2, 1)[:, :, class_index]
    t = target_onehot.flatten(2).transpose(2, 1)[:, :, class_index]
    return torch.mean((torch.sum(p * t, -1) + eps) / (torch.sum(t, -1) + eps))


def precision_multi(predict, target_onehot, class_index, eps=1e-5):
    """loss/multi_criterions.py Precision / Precision2."""
    p = predict.flatten(2).transpose(2, 1)[:, :, class_index]
    t = target_onehot.flatten(2).transpose(2, 1)[:, :, class_index]
    return torch.mean((torch.sum(p * t, -1) + eps) / (torch.sum(p, -1) + eps))


def localization_loss_multi(predict, target_onehot, eps=1e-6, mask_threshold=10):
    """loss/multi_criterions.py:232-281: foreground (1 - channel 0) H-profile distance, no factor 8."""
    p = (1 - predict[:, 0]).unsqueeze(1).float()
    t = (1 - target_onehot[:, 0]).unsqueeze(1).float()
    pp = torch.sigmoid(p.flatten(3).sum(-1) - mask_threshold)
    tp = torch.sigmoid(t.flatten(3).sum(-1) - mask_threshold)
    return _profile_distance(pp, tp, eps, 1.0)


def overlap_counts(pred_onehot: np.ndarray, target: np.ndarray) -> np.ndarray:
    """Integer statistics the metrics are functions of: int64 [C+1, H, 3] = (TP, predicted, target) per class and
    H-row, row C = foreground pseudo-class (1 - pred[0]) vs (target != 0)."""
    C, H = pred_onehot.shape[0], pred_onehot.shape[1]
    out = np.zeros((C + 1, H, 3), dtype=np.int64)
    for c in range(C + 1):
        p = (pred_onehot[0] == 0) if c == C else (pred_onehot[c] != 0)
        t = (target != 0) if c == C else (target == c)
        out[c, :, 0] = (p & t).reshape(H, -1).sum(1)
        out[c, :, 1] = p.reshape(H, -1).sum(1)
        out[c, :, 2] = t.reshape(H, -1).sum(1)
    return out
