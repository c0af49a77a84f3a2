"""GPU: the loss kernels (ltu_loss_sums, ltu_loss_sums_bwd, ltu_label_pool) and lintransunet_b200.losses on top of them
against the oracle's restatement of the reference criteria (oracle/train_step.py, pinned to loss/criterions.py through
tests/golden/train_c2_64x64x16.npz): loss terms, total and the gradients with respect to every supervised output."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ltu_oracle as O
from oracle import train_step as T
from tests.helpers import load_golden, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,C,shape", [(1, 2, (8, 8, 4)), (2, 2, (32, 32, 16)), (3, 3, (17, 9, 5)), (2, 2, (96, 96, 96))])
def test_loss_sums_and_backward(N, C, shape):
    from lintransunet_b200 import ops
    g = torch.Generator().manual_seed(3)
    p = torch.softmax(4 * torch.randn((N, C) + shape, generator=g), 1)
    p[0, 0].view(-1)[:3] = torch.tensor([0.0, 1e-7, 1.0])                 # clamp region and the exact ends
    labels = torch.randint(0, C, (N,) + shape, generator=g, dtype=torch.uint8)
    pd = p.double().requires_grad_(True)
    onehot = F.one_hot(labels.long(), C).movedim(-1, 1).double()
    s = -(1 - pd) * onehot * torch.log(torch.clamp(pd, min=1e-6))
    ref = torch.stack([pd.flatten(2).sum(-1), onehot.flatten(2).sum(-1), (pd * onehot).flatten(2).sum(-1), s.flatten(2).sum(-1)], -1)
    got = ops.loss_sums(p.cuda(), labels.cuda())
    assert got.shape == (N, C, 4) and rel_err(got, ref.detach()) < 2e-6
    assert torch.equal(got, ops.loss_sums(p.cuda(), labels.cuda()))       # fixed-order sums
    gs = torch.randn(N, C, 4, generator=g)
    ref.backward(gs.double())
    dp = ops.loss_sums_bwd(p.cuda(), labels.cuda(), gs.cuda())
    assert rel_err(dp, pd.grad) < 1e-5


def test_label_pool_is_max_pool3d():
    from lintransunet_b200 import ops
    lab = (torch.rand(2, 32, 64, 16, generator=torch.Generator().manual_seed(4)) > 0.9).to(torch.uint8)
    for k in ((2, 2, 1), (2, 2, 2)):
        ref = F.max_pool3d(lab.float().unsqueeze(1), kernel_size=k, stride=k).squeeze(1).to(torch.uint8)
        assert torch.equal(ops.label_pool(lab.cuda(), k).cpu(), ref)
    with pytest.raises(RuntimeError):
        ops.label_pool(lab.cuda(), (3, 2, 1))


def test_deep_supervision_loss_matches_reference_terms_and_oracle_gradients():
    from lintransunet_b200 import losses
    g = load_golden("train_c2_64x64x16.npz")
    cfg = O.UnetConfig(dim_output=2)
    sd = O.make_state_dict(cfg, seed=0)
    x = O.make_input((1, 1, 64, 64, 16), seed=1, blob=True)
    masks = torch.from_numpy(g["masks"]).long()
    with torch.no_grad():
        out = O.mask_trans_unet_forward(x, sd, cfg)
    probs = out["probs"].cuda().requires_grad_(True)
    mlist = [m.cuda().requires_grad_(True) for m in out["mask_list"]]
    total, terms = losses.deep_supervision_loss(probs, mlist, masks.cuda())
    got = np.asarray([[float(v.detach()) for v in row] for row in terms])
    np.testing.assert_allclose(got, g["terms"], rtol=2e-5, atol=2e-6)       # the unmodified reference's ten loss terms
    assert abs(float(total.detach()) - float(g["total"])) <= 2e-5 * abs(float(g["total"]))
    grads = torch.autograd.grad(total, [probs] + mlist)
    probs2 = out["probs"].double().requires_grad_(True)
    mlist2 = [m.double().requires_grad_(True) for m in out["mask_list"]]
    total2, _ = T.train_loss(probs2, mlist2, masks)
    grads2 = torch.autograd.grad(total2, [probs2] + mlist2)
    for a, b in zip(grads, grads2):
        assert rel_err(a, b) < 2e-5
