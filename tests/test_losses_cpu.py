"""CPU: the scalar algebra of lintransunet_b200/losses.py (the reference's deep-supervision loss recipe on the product side,
written on the four per-(sample, class) sums of ltu_loss_sums) reproduces the loss terms the unmodified reference computed
(tests/golden/train_c2_64x64x16.npz, tools/make_golden_train.py) and gives the same starting gradients as the oracle's
restatement.  The two kernels are replaced by their torch definitions here (no GPU); tests/test_losses_gpu.py checks the
kernels themselves."""
import numpy as np
import torch

from oracle import ltu_oracle as O
from oracle import train_step as T
from tests.helpers import load_golden
from tests.test_backward_composition_cpu import _install_loss_standins


def test_product_loss_matches_reference_terms_and_oracle_gradients(monkeypatch):
    from lintransunet_b200 import losses
    _install_loss_standins(monkeypatch)
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
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-10)
