"""Deep-supervision training loss of the reference, the scalar its train step backpropagates:
``train3D.py:85-93,:139-152`` (criteria and weights) + ``utils/utils_3D_embed_full.py:63-86`` (the loop over the final
probabilities and the four mask-head outputs against max-pooled labels) + ``loss/criterions.py`` (CrossEntroLoss :696-718,
DiceClassLoss :35-69, BalanceDiceLoss :416-443), binary models.

All three criteria depend on the probabilities only through four sums per (sample, class) -- sum p, sum onehot,
sum p*onehot and sum -(1-p)*onehot*log(clamp(p)) -- so each supervised output costs ONE pass of ``ltu_loss_sums`` over its
probabilities (and one pass of ``ltu_loss_sums_bwd`` in the backward) instead of the reference's flatten / transpose /
stack / clamp / log temporaries; the label pyramid is ``ltu_label_pool`` on uint8 labels.  What remains in torch is scalar
algebra on the [N, C, 4] sums, which autograd differentiates.  CUDA tensors only, like every op of this package.  Values
are pinned to the unmodified reference classes through tests/golden/train_c2_64x64x16.npz (tests/test_losses_gpu.py; the
CPU test runs the same algebra on torch stand-ins of the two kernels).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
from torch import Tensor

from . import ops

__all__ = ["cross_entro_loss", "dice_class_loss", "balance_dice_loss", "deep_supervision_loss", "level_sums", "WEIGHT_LIST"]

WEIGHT_LIST = (0.05, 0.05, 0.1, 0.1, 1.0)            # train3D.py:91-93


class _LossSums(torch.autograd.Function):
    """sums = ltu_loss_sums(p, labels), differentiable in p (ltu_loss_sums_bwd)."""

    @staticmethod
    def forward(ctx, p: Tensor, labels: Tensor) -> Tensor:
        p = p.contiguous()
        ctx.save_for_backward(p, labels)
        return ops.loss_sums(p, labels)

    @staticmethod
    def backward(ctx, g: Tensor):
        p, labels = ctx.saved_tensors
        return ops.loss_sums_bwd(p, labels, g.contiguous().to(p.dtype)), None


def level_sums(predict: Tensor, labels_u8: Tensor) -> Tuple[Tensor, int]:
    """predict fp32 [N, C, ...] probabilities, labels uint8 [N, ...] -> (sums [N, C, 4] = (A, T, X, S), voxels V)."""
    n, c = predict.shape[0], predict.shape[1]
    return _LossSums.apply(predict.float(), labels_u8), predict.numel() // (n * c)


def cross_entro_loss(sums: Tensor, voxels: int, eps: float = 1e-5) -> Tensor:
    """loss/criterions.py:701-718: mean over (n, v, c) of -w[n,c] (1-p) onehot log(clamp(p)), w = (V - (sum p + eps)) / V."""
    a, s = sums[..., 0], sums[..., 3]
    w = (voxels - (a + eps)) / voxels
    return torch.sum(w * s) / (sums.shape[0] * sums.shape[1] * voxels)


def dice_class_loss(sums: Tensor, class_index: int = 1, eps: float = 1e-9) -> Tensor:
    """loss/criterions.py:46-69: 1 - mean_n (2 sum(p_c t) + eps) / (sum(p_c + t) + eps) for class 1."""
    a, t, x = sums[:, class_index, 0], sums[:, class_index, 1], sums[:, class_index, 2]
    return 1 - torch.mean((2 * x + eps) / (a + t + eps))


def balance_dice_loss(sums: Tensor, eps: float = 1e-5) -> Tensor:
    """loss/criterions.py:424-443: class weights 1 / (sum onehot + eps)^2."""
    a, t, x = sums[..., 0], sums[..., 1], sums[..., 2]
    cw = 1 / (t + eps) ** 2
    cross = 2 * torch.sum(x * cw, dim=1) + eps
    total = torch.sum((a + t) * cw, dim=1) + eps
    return 1 - torch.mean(cross / total)


def deep_supervision_loss(probs: Tensor, mask_list: Sequence[Tensor], masks: Tensor,
                          weights: Sequence[float] = WEIGHT_LIST) -> Tuple[Tensor, List[List[Tensor]]]:
    """probs [B,2,H,W,D] (final softmax), mask_list = the four mask-head outputs, coarse to fine, masks [B,1,H,W,D] in
    {0,1}.  Returns (total, [[CE, Dice] per output in the reference's loop order: final, finest head, ..., coarsest])."""
    n = len(weights)
    B, _, H, W, D = masks.shape
    lab = masks.reshape(B, H, W, D).to(torch.uint8).contiguous()
    temp = ops.label_pool(lab, (2, 2, 1))                # utils_3D_embed_full.py:65
    terms: List[List[Tensor]] = []
    for k in range(n):
        level = n - 1 - k                              # criterions[-k-1]: (CE, BalanceDice) below level n-2, else (CE, DiceClass)
        if k == 0:
            sums, vox = level_sums(probs, lab)
        else:
            sums, vox = level_sums(mask_list[-k], temp)
            temp = ops.label_pool(temp, (2, 2, 2) if k % 2 == 0 else (2, 2, 1))       # :76-79
        dice = balance_dice_loss(sums) if level < n - 2 else dice_class_loss(sums)
        terms.append([cross_entro_loss(sums, vox), dice])
    total = sum(sum(t) * w for t, w in zip(terms, weights))
    return total, terms
