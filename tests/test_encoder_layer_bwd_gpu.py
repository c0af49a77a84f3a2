"""GPU: backward of the transformer glue kernels (ltu_add_layernorm_bwd, ltu_gelu_bwd) and of one whole
SelfAttentionLayer (lintransunet_b200/backward.py) against fp64 autograd through the oracle
(oracle/ltu_oracle.py::encoder_layer = model/trans_block.py:203-211).  SURVEY 8f-1, first slice."""
import pytest
import torch
import torch.nn.functional as F

from oracle import ltu_oracle as O
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,C", [(5, 128), (1000, 256), (4097, 128)])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 1e-2)])
def test_add_layernorm_bwd(rows, C, dtype, tol):
    from lintransunet_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(rows + C)
    x, r, dy = (torch.randn(rows, C, device="cuda", generator=g).to(dtype) for _ in range(3))
    gamma = torch.rand(C, device="cuda", generator=g) + 0.5
    beta = torch.randn(C, device="cuda", generator=g)
    xd, rd, gd, bd = (t.double().clone().requires_grad_(True) for t in (x, r, gamma, beta))
    F.layer_norm(xd + rd, (C,), gd, bd, eps=1e-6).backward(dy.double())
    dz, dgamma, dbeta = ops.add_layernorm_bwd(x, r, dy, gamma, 1e-6)
    assert dz.dtype == dtype and dgamma.dtype == torch.float32
    assert torch.equal(xd.grad, rd.grad)
    assert rel_err(dz, xd.grad) <= tol
    assert rel_err(dgamma, gd.grad) <= 2e-5 and rel_err(dbeta, bd.grad) <= 2e-5     # fp32 sums of the same inputs
    dz2, dg2, db2 = ops.add_layernorm_bwd(x, r, dy, gamma, 1e-6)
    assert torch.equal(dz, dz2) and torch.equal(dgamma, dg2) and torch.equal(dbeta, db2)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-6), (torch.bfloat16, 8e-3)])
def test_gelu_bwd(dtype, tol):
    from lintransunet_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    x = (torch.randn(333, 256, device="cuda", generator=g) * 2.5).to(dtype)
    dy = torch.randn(333, 256, device="cuda", generator=g).to(dtype)
    xd = x.double().clone().requires_grad_(True)
    F.gelu(xd).backward(dy.double())
    assert rel_err(ops.gelu_bwd(x, dy), xd.grad) <= tol
    y = ops.gelu(x)
    assert y.data_ptr() != x.data_ptr() and rel_err(y, F.gelu(x.double())) <= max(tol, 1e-6) * 4


@pytest.mark.parametrize("d_model,nhead,B,N", [(128, 4, 2, 700), (256, 8, 1, 1234), (256, 8, 3, 64)])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 3e-4), (torch.bfloat16, 4e-2)])
def test_encoder_layer_backward_matches_oracle_autograd(d_model, nhead, B, N, dtype, tol):
    from lintransunet_b200.backward import encoder_layer_backward, encoder_layer_train
    from lintransunet_b200.unet import SelfAttentionLayer
    torch.manual_seed(d_model + N)
    layer = SelfAttentionLayer(d_model, nhead).cuda()
    with torch.no_grad():
        for name, p in layer.named_parameters():              # non-trivial LayerNorm parameters and biases
            if "layer_norm" in name or name.endswith("bias"):
                p.add_(0.3 * torch.randn_like(p))
    x = torch.randn(B, N, d_model, device="cuda").to(dtype)
    dout = torch.randn(B, N, d_model, device="cuda").to(dtype)
    # reference: fp64 autograd through the oracle's layer on the same (rounded) inputs
    sd = {f"L.{k}": v.detach().double().clone().requires_grad_(True) for k, v in layer.state_dict().items()}
    xd = x.double().clone().requires_grad_(True)
    yd = O.encoder_layer(xd, sd, "L", nhead)
    yd.backward(dout.double())
    y, saved = encoder_layer_train(x, layer)
    dx, grads = encoder_layer_backward(dout, saved)
    assert rel_err(y, yd.detach()) <= tol
    e_in = rel_err(dx, xd.grad)
    assert sorted(grads) == sorted(k[2:] for k in sd)
    worst, worst_name = 0.0, ""
    for name, gr in grads.items():
        assert gr.dtype == torch.float32 and gr.shape == sd["L." + name].shape
        if name == "self_attn.linears.1.bias":
            # mathematically zero (softmax over the tokens is invariant to a per-column shift of K): only noise on
            # both sides, measured against the scale of the K-projection weight gradient
            assert float(gr.abs().max()) <= tol * float(grads["self_attn.linears.1.weight"].abs().max())
            continue
        e = rel_err(gr, sd["L." + name].grad)
        if e > worst:
            worst, worst_name = e, name
    print(f"\n[encoder layer bwd {dtype} C={d_model} B={B} N={N}] dx rel err {e_in:.2e}, worst parameter gradient "
          f"{worst:.2e} ({worst_name})")
    assert e_in <= tol and worst <= tol


@pytest.mark.parametrize("shape,C", [((2, 5, 4, 6), 128), ((1, 9, 7, 8), 256), ((2, 3, 3, 3), 128)])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 1e-2)])
def test_posenc_backward(shape, C, dtype, tol):
    """Conv3dPosEmbedding backward: dx through the forward kernel with reversed taps, dw / dbias by ltu_posenc_wgrad."""
    from lintransunet_b200 import ops
    from lintransunet_b200.unet import Conv3dPosEmbedding, _pos_w
    B, H, W, D = shape
    torch.manual_seed(H * W + C)
    pe = Conv3dPosEmbedding(C).cuda()
    w27, pb = _pos_w(pe)
    x = torch.randn(B, H, W, D, C, device="cuda").to(dtype)
    dy = torch.randn(B, H, W, D, C, device="cuda").to(dtype)
    xd = x.double().permute(0, 4, 1, 2, 3).clone().requires_grad_(True)           # oracle layout [B,C,H,W,D]
    wd = pe.proj.weight.detach().double().clone().requires_grad_(True)
    bd = pe.proj.bias.detach().double().clone().requires_grad_(True)
    O.pos_embedding(xd, wd, bd).backward(dy.double().permute(0, 4, 1, 2, 3))
    dx, dw27, db = ops.posenc_dwconv3_bwd(x, dy, w27)
    dw = dw27.reshape(3, 3, 3, C).permute(3, 2, 0, 1).unsqueeze(1)               # -> [C,1,kd,kh,kw]
    assert rel_err(dx, xd.grad.permute(0, 2, 3, 4, 1)) <= tol
    assert rel_err(dw, wd.grad) <= 2e-5 and rel_err(db, bd.grad) <= 2e-5
    dx2, dw2, db2 = ops.posenc_dwconv3_bwd(x, dy, w27)
    assert torch.equal(dw27, dw2) and torch.equal(db, db2)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-3), (torch.bfloat16, 8e-2)])
def test_transformer_stack_backward_matches_oracle_autograd(dtype, tol):
    """Eight layers + the positional conv (PosAttention3DBlock / EmbedAttention3DBlock stack) end to end."""
    from lintransunet_b200.backward import transformer_stack_backward, transformer_stack_train
    from lintransunet_b200.unet import EmbedAttention3DBlock
    torch.manual_seed(11)
    C, nhead, n_layers = 128, 4, 8
    blk = EmbedAttention3DBlock(32, C, nhead, n_layers).cuda()
    B, H, W, D = 2, 6, 5, 8
    x = torch.randn(B, H, W, D, C, device="cuda").to(dtype)
    dout = torch.randn(B, H, W, D, C, device="cuda").to(dtype)
    sd = {f"T.{k}": v.detach().double().clone().requires_grad_(True) for k, v in blk.state_dict().items()}
    xd = x.double().permute(0, 4, 1, 2, 3).clone().requires_grad_(True)
    yd = O.transformer_stack(xd, sd, "T", nhead, sd["T.pos_encoder.proj.weight"], sd["T.pos_encoder.proj.bias"], n_layers)
    yd.backward(dout.double().permute(0, 4, 1, 2, 3))
    y, saved = transformer_stack_train(x, blk.layers, blk.pos_encoder)
    dx, grads = transformer_stack_backward(dout, saved)
    assert rel_err(y, yd.detach().permute(0, 2, 3, 4, 1)) <= tol
    e_in = rel_err(dx, xd.grad.permute(0, 2, 3, 4, 1))
    worst, worst_name = 0.0, ""
    for name, gr in grads.items():
        ref = sd["T." + name.replace("pos.", "pos_encoder.")].grad
        assert gr.shape == ref.shape, name
        if name.endswith("self_attn.linears.1.bias"):                             # mathematically zero: noise only
            assert float(gr.abs().max()) <= tol * float(grads[name[:-4] + "weight"].abs().max())
            continue
        e = rel_err(gr, ref)
        if e > worst:
            worst, worst_name = e, name
    assert len(grads) == n_layers * 16 + 2
    print(f"\n[transformer stack bwd {dtype}] dx rel err {e_in:.2e}, worst parameter gradient {worst:.2e} ({worst_name})")
    assert e_in <= tol and worst <= tol


@pytest.mark.parametrize("shape,C,act", [((2, 6, 5, 7), 16, 1), ((1, 9, 8, 10), 64, 1), ((3, 4, 4, 4), 256, 1), ((2, 12, 10, 9), 32, 0)])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 5e-5), (torch.bfloat16, 1.5e-2)])
def test_instnorm_lrelu_backward(shape, C, act, dtype, tol):
    """Backward of LeakyReLU(InstanceNorm3d(x)) (model/Unet_3Dblock.py:325-336) on the raw conv output."""
    from lintransunet_b200 import ops
    B, H, W, D = shape
    g = torch.Generator(device="cuda").manual_seed(H * W * D + C)
    x = (torch.randn(B, H, W, D, C, device="cuda", generator=g) * 1.7 + 0.3).to(dtype)
    dy = torch.randn(B, H, W, D, C, device="cuda", generator=g).to(dtype)
    xd = x.double().permute(0, 4, 1, 2, 3).clone().requires_grad_(True)
    y = F.instance_norm(xd, eps=1e-5)
    if act:
        y = F.leaky_relu(y, 0.01)
    y.backward(dy.double().permute(0, 4, 1, 2, 3))
    stats = ops.chan_stats(x, 1e-5)                                   # the forward's (mean, rstd)
    dx = ops.instnorm_bwd(x, stats, dy, ops.ACT_LRELU if act else ops.ACT_NONE)
    assert dx.dtype == dtype and rel_err(dx, xd.grad.permute(0, 2, 3, 4, 1)) <= tol
    assert torch.equal(dx, ops.instnorm_bwd(x, stats, dy, ops.ACT_LRELU if act else ops.ACT_NONE))
