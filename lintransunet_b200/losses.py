"""Deep-supervision training loss of the reference, the scalar its train step backpropagates:
``train3D.py:85-93,:139-152`` (criteria and weights) + ``utils/utils_3D_embed_full.py:63-86`` (the loop over the final
probabilities and the four mask-head outputs against max-pooled labels) + ``loss/criterions.py`` (CrossEntroLoss :696-718,
DiceClassLoss :35-69, BalanceDiceLoss :416-443), binary models.

Plain differentiable torch ops on tiny reductions (a stop-gap, not kernels): they run on the device the tensors live on,
and ``torch.autograd.grad`` of the returned scalar with respect to ``probs`` / ``mask_list`` is the starting gradient of
the native backward (``ltu_head_d2s_softmax_bwd``, ``ltu_mask_softmax_bwd``).  Values are pinned to the unmodified reference
classes through tests/golden/train_c2_64x64x16.npz (tests/test_losses_cpu.py).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

__all__ = ["cross_entro_loss", "dice_class_loss", "balance_dice_loss", "deep_supervision_loss", "WEIGHT_LIST"]

WEIGHT_LIST = (0.05, 0.05, 0.1, 0.1, 1.0)            # train3D.py:91-93


def _rows(predict: Tensor, target: Tensor) -> Tuple[Tensor, Tensor]:
    """[N,C,...] -> [N,V,C] and [N,1,...] -> [N,V] (the reference's flatten(2).transpose(2, 1))."""
    return predict.flatten(2).transpose(2, 1), target.flatten(2).transpose(2, 1).squeeze(2)


def cross_entro_loss(predict: Tensor, target: Tensor, eps: float = 1e-5) -> Tensor:
    p, t = _rows(predict, target)
    onehot = torch.stack([1 - t, t], dim=-1)
    weight = torch.sum(p, dim=1, keepdim=True) + eps
    total = torch.sum(onehot, dim=(1, 2), keepdim=True)
    weight = (total - weight) / total
    return torch.mean(-weight * (1 - p) * onehot * torch.log(torch.clamp(p, min=1e-6)))


def dice_class_loss(predict: Tensor, target: Tensor, class_index: int = 1, eps: float = 1e-9) -> Tensor:
    p, t = _rows(predict, target)
    cp = p[:, :, class_index]
    return 1 - torch.mean((2 * torch.sum(cp * t, -1) + eps) / (torch.sum(cp + t, -1) + eps))


def balance_dice_loss(predict: Tensor, target: Tensor, eps: float = 1e-5) -> Tensor:
    p, t = _rows(predict, target)
    onehot = torch.stack([1 - t, t], dim=-1)
    cw = 1 / (torch.sum(onehot, dim=1, keepdim=True) + eps) ** 2
    cross = 2 * torch.sum(p * onehot * cw, dim=(1, 2)) + eps
    total = torch.sum((p + onehot) * cw, dim=(1, 2)) + eps
    return 1 - torch.mean(cross / total)


def deep_supervision_loss(probs: Tensor, mask_list: Sequence[Tensor], masks: Tensor,
                          weights: Sequence[float] = WEIGHT_LIST) -> Tuple[Tensor, List[List[Tensor]]]:
    """probs [B,2,H,W,D] (final softmax), mask_list = the four mask-head outputs, coarse to fine, masks [B,1,H,W,D] in
    {0,1}.  Returns (total, [[CE, Dice] per output in the reference's loop order: final, finest head, ..., coarsest])."""
    n = len(weights)
    temp = F.max_pool3d(masks.float(), kernel_size=(2, 2, 1), stride=(2, 2, 1))
    terms: List[List[Tensor]] = []
    for k in range(n):
        level = n - 1 - k                              # criterions[-k-1]: (CE, BalanceDice) below level n-2, else (CE, DiceClass)
        dice = balance_dice_loss if level < n - 2 else dice_class_loss
        if k == 0:
            terms.append([cross_entro_loss(probs, masks.long()), dice(probs, masks.long())])
        else:
            terms.append([cross_entro_loss(mask_list[-k], temp.long()), dice(mask_list[-k], temp.long())])
            with torch.no_grad():
                ks = 2 if k % 2 == 0 else (2, 2, 1)
                temp = F.max_pool3d(temp, kernel_size=ks, stride=ks)
    total = sum(sum(t) * w for t, w in zip(terms, weights))
    return total, terms
