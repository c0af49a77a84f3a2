"""CPU restatement of the loss recipe of the reference's train step (the caller the
backward kernels will serve).  TEST INFRASTRUCTURE ONLY (see oracle/ltu_oracle.py for the rules).

`train_loss(probs, mask_list, masks)` is the scalar the reference backpropagates
(utils/utils_3D_embed_full.py:63-86 with the default criteria and weights of train3D.py:85-93,
:139-152): deep supervision of the final probabilities and of the four mask-head outputs against
max-pooled labels.  Together with `ltu_oracle.mask_trans_unet_forward` (plain differentiable torch
ops) it gives reference gradients by autograd for every parameter.

PINNED: tools/make_golden_train.py runs the unmodified reference model in train mode (dropout 0),
the unmodified loss classes and `backward()`, and stores the loss terms and a subsample of every
parameter gradient in tests/golden/train_c2_64x64x16.npz (tests/test_train_oracle_cpu.py).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

WEIGHT_LIST = (0.05, 0.05, 0.1, 0.1, 1.0)                    # train3D.py:91-93 (epoch-0 value of every schedule's default)


def _flat(predict: Tensor, target: Tensor) -> Tuple[Tensor, Tensor]:
    p = predict.flatten(2).transpose(2, 1)
    t = target.flatten(2).transpose(2, 1).squeeze(2)
    return p, t


def cross_entro_loss(predict: Tensor, target: Tensor, eps: float = 1e-5) -> Tensor:
    """loss/criterions.py:701-718 (binary: one-hot = stack(1 - t, t))."""
    p, t = _flat(predict, target)
    onehot = torch.stack([1 - t, t], dim=-1)
    logp = torch.log(torch.clamp(p, min=1e-6))
    weight = torch.sum(p, dim=1, keepdim=True) + eps
    total = torch.sum(onehot, dim=(1, 2), keepdim=True)
    weight = (total - weight) / total
    return torch.mean(-weight * (1 - p) * onehot * logp)


def dice_class_loss(predict: Tensor, target: Tensor, class_index: int = 1, eps: float = 1e-9) -> Tensor:
    """loss/criterions.py:46-69."""
    p, t = _flat(predict, target)
    cp = p[:, :, class_index]
    return 1 - torch.mean((2 * torch.sum(cp * t, -1) + eps) / (torch.sum(cp + t, -1) + eps))


def balance_dice_loss(predict: Tensor, target: Tensor, eps: float = 1e-5) -> Tensor:
    """loss/criterions.py:424-443."""
    p, t = _flat(predict, target)
    onehot = torch.stack([1 - t, t], dim=-1)
    cw = 1 / (torch.sum(onehot, dim=1, keepdim=True) + eps) ** 2
    cross = 2 * torch.sum(p * onehot * cw, dim=(1, 2)) + eps
    total = torch.sum((p + onehot) * cw, dim=(1, 2)) + eps
    return 1 - torch.mean(cross / total)


def criteria_for_level(i: int, n_levels: int = 5):
    """train3D.py:139-152: levels 0..n-3 -> (CE, BalanceDice); level n-2 and the final output -> (CE, DiceClass)."""
    return (cross_entro_loss, balance_dice_loss) if i < n_levels - 2 else (cross_entro_loss, dice_class_loss)


def train_loss(probs: Tensor, mask_list: Sequence[Tensor], masks: Tensor,
               weights: Sequence[float] = WEIGHT_LIST) -> Tuple[Tensor, List[List[Tensor]]]:
    """utils/utils_3D_embed_full.py:63-86.  probs [B,2,H,W,D], mask_list = the 4 mask-head outputs (coarse to fine),
    masks [B,1,H,W,D] in {0,1}.  Returns (total, per-output [CE, Dice] terms in the reference's loop order)."""
    n = len(weights)
    temp = F.max_pool3d(masks.float(), kernel_size=(2, 2, 1), stride=(2, 2, 1))
    terms: List[List[Tensor]] = []
    for k in range(n):
        crit = criteria_for_level(n - 1 - k, n)              # criterions[-k-1]
        if k == 0:
            terms.append([c(probs, masks.long()) for c in crit])
        else:
            terms.append([c(mask_list[-k], temp.long()) for c in crit])
            with torch.no_grad():
                ks = 2 if k % 2 == 0 else (2, 2, 1)
                temp = F.max_pool3d(temp, kernel_size=ks, stride=ks)
    total = sum(sum(t) * w for t, w in zip(terms, weights))
    return total, terms
