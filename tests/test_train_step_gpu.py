"""GPU: one training step's loss and gradients from the native backward (lintransunet_b200.backward.model_loss_and_gradients)
against the gradients the unmodified reference produced on the same weights, input and labels
(tests/golden/train_c2_64x64x16.npz, tools/make_golden_train.py).

SKIPPED until its first GPU run: the decoder-loop composition was written after round 1's GPU budget was spent.  Expectation
when enabled: loss terms within a few percent (bf16 forward); gradients are compared in the relative L2 norm with a loose bound,
because against an exact fp32 reference a bf16 forward flips the LeakyReLU sign of near-zero activations (5-12 % per layer, see
tests/test_conv_bwd_gpu.py) and may move a ROI box by a pixel."""
import numpy as np
import pytest
import torch

from oracle import ltu_oracle as O
from tests.helpers import load_golden

import os

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("LTU_RUN_TRAIN_STEP") != "1",
                                 reason="decoder-loop composition not yet run on a GPU (round 1 budget spent); "
                                        "LTU_RUN_TRAIN_STEP=1 runs it")]


def test_train_step_gradients_against_the_reference():
    from lintransunet_b200 import MaskTransUnet
    from lintransunet_b200.backward import model_loss_and_gradients
    g = load_golden("train_c2_64x64x16.npz")
    cfg = O.UnetConfig(dim_output=2)
    sd = O.make_state_dict(cfg, seed=0)
    m = MaskTransUnet(list(cfg.num_layers), list(cfg.roi_size_list), list(cfg.is_roi_list), 1, 2, dropout=0.0)
    m.load_state_dict(sd)
    m.cuda()
    x = O.make_input((1, 1, 64, 64, 16), seed=1, blob=True).cuda()
    masks = torch.from_numpy(g["masks"]).long().cuda()
    total, terms, grads = model_loss_and_gradients(m, x, masks)
    got = np.asarray([[float(v.detach()) for v in row] for row in terms])
    print("\n[train step] loss terms", got.round(4).tolist(), "reference", g["terms"].round(4).tolist())
    np.testing.assert_allclose(got, g["terms"], rtol=5e-2, atol=5e-3)
    names = [str(n) for n in g["grad_names"]]
    assert sorted(grads) == sorted(names)
    ref_norm = dict(zip(names, (float(v) for v in g["grad_norms"])))
    floor = 1e-3 * max(ref_norm.values())
    worst = 0.0
    for name in names:
        if ref_norm[name] < floor:
            continue
        e = abs(float(grads[name].double().norm()) - ref_norm[name]) / ref_norm[name]
        worst = max(worst, e)
    print(f"[train step] worst relative gradient-norm deviation above the floor: {worst:.2e}")
    assert worst <= 0.35
