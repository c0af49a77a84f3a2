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
    # element-wise: the stored 64-element subsample of every gradient (tools/make_golden_train.py::gsub) -- catches a sign
    # flip, a transposed dW layout or gradients routed to the wrong same-shape parameter, which keep the norm
    cos_worst, n_checked = 1.0, 0
    for name in names:
        if ref_norm[name] < 10 * floor:
            continue
        ref = g["g:" + name].astype(np.float64)
        f = grads[name].detach().reshape(-1)
        step = max(1, f.numel() // 64)
        got_sub = f[::step][:64].double().cpu().numpy()
        if np.linalg.norm(ref) < 1e-3 * ref_norm[name]:
            continue                                    # subsample carries no signal
        cos = float(np.dot(got_sub, ref) / (np.linalg.norm(got_sub) * np.linalg.norm(ref) + 1e-300))
        cos_worst = min(cos_worst, cos)
        n_checked += 1
        assert cos > 0.8, (name, cos)
    print(f"[train step] worst cosine similarity of a gradient subsample with the reference's: {cos_worst:.4f} over {n_checked} tensors")
    assert n_checked > 50


def _model(dropout=0.0, seed=0):
    from lintransunet_b200 import MaskTransUnet
    cfg = O.UnetConfig(dim_output=2)
    m = MaskTransUnet(list(cfg.num_layers), list(cfg.roi_size_list), list(cfg.is_roi_list), 1, 2, dropout=dropout)
    m.load_state_dict(O.make_state_dict(cfg, seed=seed))
    return m.cuda()


def test_autograd_wiring_fills_param_grads():
    """loss.backward() through unet._NativeTrainFunction (the DEFAULT training path: model.train(), autocast, grad enabled)
    fills p.grad with exactly what backward.model_loss_and_gradients returns, leaves the dead parameters without a
    gradient, and train.train_step moves the weights."""
    from lintransunet_b200 import losses
    from lintransunet_b200.backward import model_loss_and_gradients
    from lintransunet_b200.train import train_step
    g = load_golden("train_c2_64x64x16.npz")
    m = _model()
    assert m.native_backward
    x = O.make_input((1, 1, 64, 64, 16), seed=1, blob=True).cuda()
    masks = torch.from_numpy(g["masks"]).long().cuda()
    total_ref, _, grads = model_loss_and_gradients(m, x, masks)
    m.train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        probs, mask_list = m(x)
    assert probs.requires_grad and len(mask_list) == 4
    total, _ = losses.deep_supervision_loss(probs, mask_list, masks)
    assert abs(float(total) - float(total_ref)) <= 1e-6 * abs(float(total_ref))
    total.backward()
    dead = {str(n) for n in g["dead_names"]}
    for name, p in m.named_parameters():
        if name in dead:
            assert p.grad is None, name
            continue
        assert p.grad is not None and p.grad.shape == p.shape, name
        assert torch.equal(p.grad, grads[name].to(p.grad.dtype)), name        # same kernels, fixed-order reductions
    # the reference's fp16 autocast + GradScaler (utils/utils_3D_embed_full.py:64,:87-92) also drives it
    m.zero_grad()
    scaler = torch.amp.GradScaler("cuda", init_scale=1024.0)
    with torch.autocast("cuda", dtype=torch.float16):
        probs, mask_list = m(x)
        total2, _ = losses.deep_supervision_loss(probs, mask_list, masks)
    scaler.scale(total2).backward()
    pname = "decode.final_block.weight"
    pg = dict(m.named_parameters())[pname].grad
    assert torch.allclose(pg / 1024.0, grads[pname], rtol=2e-2, atol=1e-3 * float(grads[pname].abs().max()))
    m.zero_grad()
    opt = torch.optim.SGD(m.parameters(), lr=1e-2)
    w0 = m.decode.final_block.weight.detach().clone()
    loss, terms = train_step(m, opt, x, masks)
    assert np.isfinite(loss) and len(terms) == 5
    assert not torch.equal(w0, m.decode.final_block.weight.detach())
    # the packed-weight cache follows the optimizer's in-place update: a second step sees the new weights
    loss2, _ = train_step(m, opt, x, masks)
    assert loss2 != loss


def test_training_forward_needs_autocast():
    m = _model()
    m.train()
    x = O.make_input((1, 1, 64, 64, 16), seed=1, blob=True).cuda()
    with pytest.raises(NotImplementedError):
        m(x)                                  # fp32 training: there is no fp32 backward
    with torch.no_grad():
        probs, mask_list = m(x)               # dropout-free, no grad: the fp32 inference kernels, train-mode outputs
    assert probs.shape == (1, 2, 64, 64, 16) and len(mask_list) == 4
