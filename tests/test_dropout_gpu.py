"""GPU: training-mode dropout (ltu_dropout, backward.DropoutState): the op's distribution and determinism, the gradient of
an encoder layer WITH dropout against fp64 autograd using the very same masks, and the model-level drop-in behaviour the
reference's train3D*.py scripts rely on (MaskTransUnet(dropout=0.3).train() under autocast; reference sites:
model/trans_block.py:96,:205,:208,:209, model/Unet_3Dblock.py:339,:382,:429,:556)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ltu_oracle as O
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


def _ops():
    from lintransunet_b200 import ops
    return ops


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("p", [0.3, 0.1])
def test_dropout_elementwise(dtype, p):
    ops = _ops()
    x = (torch.rand(4, 33, 17, 8, 64, device="cuda") + 0.5).to(dtype)
    y = ops.dropout(x, p, seed=1234, offset=40)
    keep = y != 0
    frac = float(keep.float().mean())
    n = x.numel()
    assert abs(frac - (1 - p)) < 5 * np.sqrt(p * (1 - p) / n) + 1e-4, frac
    expect = (x.float() / (1 - p)).to(dtype)
    assert torch.equal(y[keep], expect[keep])
    assert torch.equal(y, ops.dropout(x, p, seed=1234, offset=40))               # pure function of (seed, offset, index)
    assert not torch.equal(y != 0, ops.dropout(x, p, seed=1235, offset=40) != 0)
    y2 = ops.dropout(x, p, seed=1234, offset=40 + ops.dropout_counters(x))       # the next site's draws are independent
    agree = float(((y2 != 0) == keep).float().mean())
    assert abs(agree - (p * p + (1 - p) * (1 - p))) < 5e-3
    # no structure along any axis: keep rate per channel and per leading slice
    assert float(keep.float().mean(dim=(0, 1, 2, 3)).std()) < 3 * np.sqrt(p * (1 - p) / (n / 64))
    # in place == out of place; the backward is the same call on the gradient
    z = x.clone()
    assert ops.dropout(z, p, 1234, 40, inplace=True) is z and torch.equal(z, y)
    dy = torch.randn_like(x)
    dx = ops.dropout(dy, p, 1234, 40)
    assert torch.equal(dx != 0, keep & (dy != 0))


def test_dropout_channelwise_is_dropout3d():
    ops = _ops()
    B, C = 6, 256
    x = (torch.rand(B, 5, 6, 4, C, device="cuda") + 0.5).bfloat16()
    y = ops.dropout(x, 0.3, seed=7, offset=0, channelwise=True)
    keep = (y != 0).reshape(B, -1, C)
    assert torch.equal(keep.all(dim=1) | (~keep).all(dim=1), torch.ones(B, C, dtype=torch.bool, device="cuda"))   # whole channels
    per = keep[:, 0, :].float()
    assert abs(float(per.mean()) - 0.7) < 5 * np.sqrt(0.21 / (B * C))
    assert not torch.equal(per[0], per[1])                                        # a different draw per sample
    assert ops.dropout_counters(x, True) == B * C // 4
    with pytest.raises(RuntimeError):
        ops.dropout(x, 1.0, 1, 0)


def test_encoder_layer_gradients_with_dropout_match_autograd_on_the_same_masks():
    """fp32 layer, p = 0.3: forward and all 17 gradients against fp64 autograd of the reference's layer
    (trans_block.py:203-211) with dropout1 / dropout / dropout2 replaced by the masks ltu_dropout draws."""
    from lintransunet_b200.backward import DropoutState, encoder_layer_train, encoder_layer_backward
    from lintransunet_b200.unet import SelfAttentionLayer
    ops = _ops()
    torch.manual_seed(0)
    B, N, C, h, p = 2, 200, 128, 4, 0.3
    layer = SelfAttentionLayer(C, h).cuda()
    x = torch.randn(B, N, C, device="cuda")
    dout = torch.randn(B, N, C, device="cuda")
    drop = DropoutState(p, seed=99, offset=8)
    y, saved = encoder_layer_train(x, layer, drop)
    dx, grads = encoder_layer_backward(dout, saved)
    # the masks of the three sites, in draw order: o [B,N,C], gelu output [B,N,2C], linear2 output [B,N,C]
    off = 8
    masks = []
    for width in (C, 2 * C, C):
        ones = torch.ones(B, N, width, device="cuda")
        masks.append(ops.dropout(ones, p, 99, off).double())                      # 0 or 1 / (1 - p)
        off += ops.dropout_counters(ones)
    assert drop.offset == off
    ref_layer = SelfAttentionLayer(C, h).cuda().double()
    ref_layer.load_state_dict(layer.state_dict())
    sd = {"L." + k: v for k, v in ref_layer.named_parameters()}
    xd = x.double().requires_grad_(True)
    a = O.multihead_attention(xd, sd, "L.self_attn", h)
    t1 = F.layer_norm(xd + a * masks[0], (C,), sd["L.layer_norm1.weight"], sd["L.layer_norm1.bias"], eps=1e-6)
    f = F.linear(F.gelu(F.linear(t1, sd["L.linear1.weight"], sd["L.linear1.bias"])) * masks[1],
                 sd["L.linear2.weight"], sd["L.linear2.bias"])
    ref = F.layer_norm(t1 + f * masks[2], (C,), sd["L.layer_norm2.weight"], sd["L.layer_norm2.bias"], eps=1e-6)
    assert rel_err(y, ref.detach()) < 1e-4
    ref.backward(dout.double())
    assert rel_err(dx, xd.grad) < 3e-4
    for k, v in ref_layer.named_parameters():
        if float(v.grad.abs().max()) < 1e-9:            # the key bias: softmax over the tokens is invariant to it
            assert float(grads[k].abs().max()) < 1e-5, k
            continue
        assert rel_err(grads[k], v.grad) < 3e-4, k


def _model(dropout):
    from lintransunet_b200 import MaskTransUnet
    cfg = O.UnetConfig(dim_output=2)
    m = MaskTransUnet(list(cfg.num_layers), list(cfg.roi_size_list), list(cfg.is_roi_list), 1, 2, dropout=dropout)
    m.load_state_dict(O.make_state_dict(cfg, seed=0))
    return m.cuda()


def test_model_train_mode_with_reference_default_dropout():
    """The reference constructs MaskTransUnet with dropout=0.3 and trains under autocast (train3D.py:119,
    utils/utils_3D_embed_full.py:63-91): train() draws fresh masks every forward, is reproducible under
    torch.manual_seed, backpropagates through them, and eval() is untouched by the dropout value."""
    from lintransunet_b200 import losses
    m = _model(0.3)
    x = O.make_input((1, 1, 64, 64, 16), seed=1, blob=True).cuda()
    masks = (torch.rand(1, 1, 64, 64, 16, device="cuda") > 0.7).long()
    m.train()
    torch.manual_seed(5)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        p1, ml1 = m(x)
        p2, _ = m(x)
    torch.manual_seed(5)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        p3, ml3 = m(x)
    assert not torch.equal(p1, p2)                      # consecutive forwards use different masks
    assert torch.equal(p1, p3) and all(torch.equal(a, b) for a, b in zip(ml1, ml3))
    assert torch.isfinite(p1).all() and abs(float(p1.sum(1).mean()) - 1.0) < 1e-3
    total, _ = losses.deep_supervision_loss(p3, ml3, masks)
    total.backward()
    got = [n for n, q in m.named_parameters() if q.grad is not None]
    assert len(got) > 500 and all(torch.isfinite(q.grad).all() for q in m.parameters() if q.grad is not None)
    # no-grad training forward (e.g. a train-mode validation pass) also applies dropout
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        p4, _ = m(x)
    assert not torch.equal(p4, p1)
    # eval: identical to a dropout-free model
    m.eval()
    m0 = _model(0.0).eval()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert torch.equal(m(x), m0(x))


def test_dropout_gradient_is_the_derivative_of_the_masked_forward():
    """Directional-derivative check of the WHOLE training step with dropout 0.3: for a fixed generator state the forward
    is a deterministic function of the weights; its change along the gradient direction of the final head's bias equals
    <grad, delta> (fp32 head, so bf16 storage noise does not enter this parameter's path)."""
    from lintransunet_b200 import losses
    m = _model(0.3)
    x = O.make_input((1, 1, 64, 64, 16), seed=1, blob=True).cuda()
    masks = (torch.rand(1, 1, 64, 64, 16, device="cuda") > 0.7).long()
    m.train()

    def loss_of():
        torch.manual_seed(11)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            pr, ml = m(x)
        return losses.deep_supervision_loss(pr, ml, masks)[0]

    total = loss_of()
    total.backward()
    b = m.decode.final_block.bias
    gvec = b.grad.detach().clone()
    assert float(gvec.norm()) > 0
    eps = 1e-2 / float(gvec.norm())
    with torch.no_grad():
        b.add_(eps * gvec)
    up = float(loss_of().detach())
    with torch.no_grad():
        b.add_(-2 * eps * gvec)
    down = float(loss_of().detach())
    fd = (up - down) / (2 * eps)
    an = float((gvec * gvec).sum())
    assert abs(fd - an) <= 5e-2 * abs(an) + 1e-6, (fd, an)
