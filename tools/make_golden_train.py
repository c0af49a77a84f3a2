#!/usr/bin/env python
"""Generate tests/golden/train_c2_64x64x16.npz: one training step's loss terms and parameter
gradients from the UNMODIFIED reference (model in train mode with dropout=0.0, the loss classes of
loss/criterions.py wired as train3D.py:139-152 / utils/utils_3D_embed_full.py:63-86, autograd
backward), on the seeded weights and input the forward vectors use.  fp32 on the CPU.

Pins oracle/train_step.py + autograd through oracle/ltu_oracle.py: the gradient reference the
backward kernels will be tested against.

Usage:  python tools/make_golden_train.py [--ref /root/reference]
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ltu_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
WEIGHTS = [0.05, 0.05, 0.1, 0.1, 1.0]                      # train3D.py:91-93
SHAPE = (1, 1, 64, 64, 16)


def gsub(t: torch.Tensor, n: int = 64) -> np.ndarray:
    f = t.detach().reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step][:n].to(torch.float32).numpy().copy()


def make_masks(shape, seed):
    g = torch.Generator().manual_seed(seed)
    m = torch.zeros(shape, dtype=torch.long)
    m[:, :, 20:44, 16:40, 4:12] = 1
    flip = torch.rand(shape, generator=g) < 0.02
    return torch.where(flip, 1 - m, m)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    sys.path.insert(0, args.ref)
    from model.trans_3DUnet import get_model_dict
    from loss.criterions import get_criterions
    cfg = O.UnetConfig(dim_output=2)
    sd = O.make_state_dict(cfg, seed=0)
    m = get_model_dict("MaskTransUnet")(num_layers=list(cfg.num_layers), roi_size_list=list(cfg.roi_size_list),
                                        is_roi_list=list(cfg.is_roi_list), dim_input=1, dim_output=2, dropout=0.0)
    m.load_state_dict(sd, strict=True)
    m.train()
    x = O.make_input(SHAPE, seed=1, blob=True)
    masks = make_masks(SHAPE, seed=7)
    n = len(cfg.num_layers)
    criterions = []                                          # train3D.py:139-152
    for i in range(n):
        if i < n - 2:
            criterions.append(get_criterions(["CrossEntroLoss", "BalanceDiceLoss"]))
        else:
            criterions.append(get_criterions(["CrossEntroLoss", "DiceClassLoss"]))
    predict, roi_mask = m(x)
    temp_masks = F.max_pool3d(masks.float(), kernel_size=(2, 2, 1), stride=(2, 2, 1))   # utils_3D_embed_full.py:64-80
    loss_list = []
    for k in range(len(WEIGHTS)):
        if k == 0:
            temp_loss = [l(predict, masks.long()) for l in criterions[-k - 1].values()]
        else:
            temp_loss = [l(roi_mask[-k], temp_masks.long()) for l in criterions[-k - 1].values()]
            with torch.no_grad():
                if k % 2 == 0:
                    temp_masks = F.max_pool3d(temp_masks, kernel_size=2, stride=2)
                else:
                    temp_masks = F.max_pool3d(temp_masks, kernel_size=(2, 2, 1), stride=(2, 2, 1))
        loss_list.append(temp_loss)
    total = sum([sum(loss) * w for loss, w in zip(loss_list, WEIGHTS)])
    total.backward()
    out = {"masks": masks.numpy().astype(np.uint8), "total": np.float64(total.item()),
           "terms": np.asarray([[float(v) for v in row] for row in loss_list], dtype=np.float64)}
    names, norms = [], []
    for name, p in m.named_parameters():
        if p.grad is None:
            continue
        names.append(name)
        norms.append(float(p.grad.double().norm()))
        out["g:" + name] = gsub(p.grad)
    out["grad_names"] = np.asarray(names)
    out["grad_norms"] = np.asarray(norms, dtype=np.float64)
    dead = [name for name, p in m.named_parameters() if p.grad is None]
    out["dead_names"] = np.asarray(dead)
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, "train_c2_64x64x16.npz")
    np.savez_compressed(path, **out)
    print("total", total.item(), "terms", out["terms"].tolist())
    print(len(names), "parameters with gradients,", len(dead), "without;", os.path.getsize(path), "bytes ->", path)


if __name__ == "__main__":
    main()
