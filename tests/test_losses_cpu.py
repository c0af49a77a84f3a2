"""CPU: lintransunet_b200/losses.py (the reference's deep-supervision loss recipe on the product side) reproduces the
loss terms the unmodified reference computed (tests/golden/train_c2_64x64x16.npz, tools/make_golden_train.py) and gives
the same starting gradients as the oracle's restatement."""
import numpy as np
import torch

from lintransunet_b200 import losses
from oracle import ltu_oracle as O
from oracle import train_step as T
from tests.helpers import load_golden


def test_product_loss_matches_reference_terms_and_oracle_gradients():
    g = load_golden("train_c2_64x64x16.npz")
    cfg = O.UnetConfig(dim_output=2)
    sd = O.make_state_dict(cfg, seed=0)
    x = O.make_input((1, 1, 64, 64, 16), seed=1, blob=True)
    masks = torch.from_numpy(g["masks"]).long()
    with torch.no_grad():
        out = O.mask_trans_unet_forward(x, sd, cfg)
    probs = out["probs"].clone().requires_grad_(True)
    mlist = [m.clone().requires_grad_(True) for m in out["mask_list"]]
    total, terms = losses.deep_supervision_loss(probs, mlist, masks)
    got = np.asarray([[float(v.detach()) for v in row] for row in terms])
    np.testing.assert_allclose(got, g["terms"], rtol=2e-5, atol=2e-6)
    assert abs(float(total.detach()) - float(g["total"])) <= 2e-5 * abs(float(g["total"]))
    grads = torch.autograd.grad(total, [probs] + mlist)
    probs2 = out["probs"].clone().requires_grad_(True)
    mlist2 = [m.clone().requires_grad_(True) for m in out["mask_list"]]
    total2, _ = T.train_loss(probs2, mlist2, masks)
    grads2 = torch.autograd.grad(total2, [probs2] + mlist2)
    for a, b in zip(grads, grads2):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-12)
