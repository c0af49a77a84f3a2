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
