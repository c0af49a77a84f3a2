"""One optimisation step of the reference's training loop (utils/utils_3D_embed_full.py:63-91) on the native forward and
backward: forward under bf16 autocast, deep-supervision loss, ``backward()`` through ``unet._NativeTrainFunction``,
optimizer step every ``step_times`` calls.

Differences from the reference, both deliberate: bf16 instead of fp16 autocast, hence no ``GradScaler`` (bf16 has fp32's
exponent range); the model must be built with ``dropout=0.0`` (training-mode dropout is not implemented).  The native
backward is opt-in (``model.native_backward``) until its autograd wrapper has run on a GPU; the gradient chain it calls is
verified against the reference (tests/test_train_step_gpu.py)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

from . import losses

__all__ = ["train_step"]


def train_step(model, optimizer: torch.optim.Optimizer, batch_images: torch.Tensor, batch_masks: torch.Tensor,
               weights: Sequence[float] = losses.WEIGHT_LIST, step_times: int = 1, do_step: bool = True
               ) -> Tuple[float, List[List[float]]]:
    """batch_images fp32 [B,1,H,W,D], batch_masks [B,1,H,W,D] in {0,1}.  Returns (total loss, [[CE, Dice] per output]).
    ``step_times``: gradient-accumulation factor of the reference (the loss is divided by it); ``do_step`` False only
    accumulates."""
    model.train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        predict, roi_mask = model(batch_images)
    total, terms = losses.deep_supervision_loss(predict, roi_mask, batch_masks, weights)
    (total / step_times).backward()
    if do_step:
        optimizer.step()
        optimizer.zero_grad()
    return float(total.detach()), [[float(v.detach()) for v in row] for row in terms]
