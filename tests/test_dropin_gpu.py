"""GPU: the API-level behaviour the reference's scripts rely on besides forward / state_dict:
whole-module pickles (train3D.py:291 `torch.save(model, ...)`, inference_*.py `torch.load`) and nn.DataParallel
(train3D.py:119), including weight changes between forwards (the derived-weight cache must follow the broadcast copies)
and a training step through the replicas."""
import io

import pytest
import torch

from oracle import ltu_oracle as O

pytestmark = pytest.mark.gpu


def _model(seed=0, dropout=0.0, classes=2):
    from lintransunet_b200 import MaskTransUnet
    cfg = O.UnetConfig(dim_output=classes)
    m = MaskTransUnet(list(cfg.num_layers), list(cfg.roi_size_list), list(cfg.is_roi_list), 1, classes, dropout=dropout)
    m.load_state_dict(O.make_state_dict(cfg, seed=seed))
    return m.cuda()


def test_whole_module_pickle_round_trip():
    m = _model().eval()
    x = O.make_input((1, 1, 64, 64, 16), seed=1, blob=True).cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = m(x)                                            # fills the plan cache and captures a CUDA graph
    buf = io.BytesIO()
    torch.save(m, buf)                                      # the reference's checkpoint format: the module itself
    buf.seek(0)
    m2 = torch.load(buf, weights_only=False)
    assert type(m2) is type(m) and not m2.training and m2._graphs == {} and m2._plans == {}
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert torch.equal(m2(x), y)
    assert torch.equal(m2(x), m(x))                         # fp32 path as well
    # the import-path shim pickles to the same class
    from model.trans_3DUnet import MaskTransUnet as Shim
    assert Shim is type(m)


def test_dataparallel_single_device_wrapper():
    m = _model().eval()
    dp = torch.nn.DataParallel(m, device_ids=[0])
    x = O.make_input((2, 1, 64, 64, 16), seed=2, blob=True).cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert torch.equal(dp(x), m(x))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_dataparallel_replicas_follow_weight_changes_and_train():
    from lintransunet_b200 import losses
    m = _model(seed=0).eval()
    dp = torch.nn.DataParallel(m, device_ids=[0, 1])
    x = O.make_input((2, 1, 64, 64, 16), seed=2, blob=True).cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y_dp, y_1 = dp(x), m(x)
    assert torch.equal(y_dp, y_1)                           # a sample's result does not depend on batch composition
    # new weights: the replicas of the next forward must not reuse derived weights of the old broadcast copies
    cfg = O.UnetConfig(dim_output=2)
    m.load_state_dict(O.make_state_dict(cfg, seed=3))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y_dp2, y_12 = dp(x), m(x)
    assert torch.equal(y_dp2, y_12) and not torch.equal(y_dp2, y_dp)
    with torch.no_grad():                                   # in-place update (an optimizer step) as well
        m.decode.final_block.bias.add_(0.5)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert torch.equal(dp(x), m(x))
    # training through the replicas: gradients arrive on the wrapped module's parameters and match the single-GPU step
    masks = (torch.rand(2, 1, 64, 64, 16, device="cuda") > 0.7).long()
    dp.train()

    def step(net):
        m.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            pr, ml = net(x)
        losses.deep_supervision_loss(pr, ml, masks)[0].backward()
        return {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}

    g_dp, g_1 = step(dp), step(m)
    # nn.DataParallel's Broadcast backward hands zero gradients to parameters the replicas never used (the dead
    # pos_encoders.1-7, the unsupervised heads): extra keys are fine as long as they are exactly zero
    assert set(g_1) <= set(g_dp) and len(g_1) > 500
    assert all(float(g_dp[k].abs().max()) == 0.0 for k in set(g_dp) - set(g_1))
    worst = max(float((g_dp[k] - g_1[k]).abs().max() / g_1[k].abs().max().clamp_min(1e-12)) for k in g_1)
    print(f"\n[DataParallel] worst relative gradient difference vs the single-GPU step: {worst:.2e}")
    assert worst < 5e-2
