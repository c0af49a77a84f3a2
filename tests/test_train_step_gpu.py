"""GPU: one training step's loss and gradients from the native backward (lintransunet_b200.backward.model_loss_and_gradients:
encoder + ROIDecoder loop + deep-supervision loss, bf16 activations) against what the UNMODIFIED reference produced on the same
weights, input and labels in fp32 (tests/golden/train_c2_64x64x16.npz, tools/make_golden_train.py).

Measured on B200 (profiles/r1_train_step_gpu.log): the ten loss terms agree to 1e-3 (0.0762/0.8506 ... vs 0.0762/0.8504 ...), the worst
relative deviation of a parameter-gradient norm (over the gradients above 1e-3 of the largest) is 4.6e-2.  The bounds below are wide on
purpose: against an exact fp32 reference a bf16 forward flips the LeakyReLU sign of near-zero activations (5-12 % per layer in
the max norm, see tests/test_conv_bwd_gpu.py) and may move a ROI box by a pixel; the kernels are bit-reproducible, so the
measured values do not fluctuate."""
import numpy as np
import pytest
import torch

from oracle import ltu_oracle as O
from tests.helpers import load_golden

pytestmark = pytest.mark.gpu


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
