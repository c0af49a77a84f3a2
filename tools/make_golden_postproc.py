#!/usr/bin/env python
"""Generate tests/golden/postproc.npz: the evaluation metrics of the reference's inference scripts
(loss/criterions.py for inference_embed_attn.py, loss/multi_criterions.py for
inference_multi_classes.py) evaluated by the UNMODIFIED reference classes, imported from
/root/reference (build container only), on seeded hard predictions and labels.

The vectors pin oracle/postproc.py and lintransunet_b200.inference.metrics_from_counts
(tests/test_postproc_cpu.py) and travel to the GPU box for tests/test_postproc_gpu.py.

Usage:  python tools/make_golden_postproc.py [--ref /root/reference]
"""
from __future__ import annotations

import argparse
import importlib
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

BINARY = ["DiceClassLoss", "Recall", "Precision", "LocalizationLoss"]                    # inference_embed_attn.py:62-64
MULTI = ["DiceClassLoss0", "DiceClassLoss", "DiceClassLoss2", "Recall", "Precision", "Recall2", "Precision2",
         "LocalizationLoss"]                                                             # inference_multi_classes.py:57-59


def blobs(shape, n_classes, seed, empty_pred=False, empty_target=False):
    """Hard label maps with a few boxes per class (rows with more and with fewer than 10 voxels: the sigmoid of
    LocalizationLoss is centred on 10)."""
    g = np.random.default_rng(seed)
    H, W, D = shape

    def one():
        lab = np.zeros(shape, dtype=np.uint8)
        for c in range(1, n_classes):
            for _ in range(3):
                h0, w0, d0 = g.integers(0, H - 4), g.integers(0, W - 4), g.integers(0, D - 3)
                dh, dw, dd = g.integers(1, 9), g.integers(1, 7), g.integers(1, 5)
                lab[h0:h0 + dh, w0:w0 + dw, d0:d0 + dd] = c
        return lab
    pred, target = one(), one()
    keep = g.random(shape) < 0.5                          # overlap between prediction and label
    pred = np.where(keep, target, pred).astype(np.uint8)
    if empty_pred:
        pred[:] = 0
    if empty_target:
        target[:] = 0
    return pred, target


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    sys.path.insert(0, args.ref)
    crit = importlib.import_module("loss.criterions")
    mcrit = importlib.import_module("loss.multi_criterions")
    out = {}
    cases = [("b0", 2, (24, 20, 12), 1, {}), ("b1", 2, (40, 16, 8), 2, {}), ("b_empty_pred", 2, (24, 20, 12), 3, dict(empty_pred=True)),
             ("b_empty_target", 2, (24, 20, 12), 4, dict(empty_target=True)),
             ("m0", 3, (24, 20, 12), 5, {}), ("m1", 3, (32, 24, 16), 6, {}), ("m_empty_pred", 3, (24, 20, 12), 7, dict(empty_pred=True)),
             ("m_empty_target", 3, (24, 20, 12), 8, dict(empty_target=True))]
    for name, C, shape, seed, kw in cases:
        pred, target = blobs(shape, C, seed, **kw)
        p = torch.from_numpy(pred).long()
        predict = F.one_hot(p, C).permute(3, 0, 1, 2)[None].float()            # [1,C,H,W,D] hard one-hot (predict2)
        if C == 2:
            masks = torch.from_numpy(target).long()[None, None]                 # [1,1,H,W,D] (inference_embed_attn.py:149)
            vals = [float(l(predict, masks)) for l in crit.get_criterions(BINARY).values()]
        else:
            label = F.one_hot(torch.from_numpy(target).long(), C).permute(3, 0, 1, 2)[None]   # :128-135
            vals = [float(l(predict, label)) for l in mcrit.get_criterions(MULTI).values()]
        out[f"{name}_pred"], out[f"{name}_target"] = pred, target
        out[f"{name}_values"] = np.asarray(vals, dtype=np.float64)
        print(name, C, shape, [round(v, 6) for v in vals])
    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, "postproc.npz"), **out)
    print("wrote", os.path.join(GOLD, "postproc.npz"))


if __name__ == "__main__":
    main()
