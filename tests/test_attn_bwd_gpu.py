"""GPU: backward of the linear-attention core (ltu_attn_bwd, SURVEY 8f-1 first slice) against autograd through the
oracle's restatement of `linear_attention` (model/trans_block.py:41-67) in fp64."""
import pytest
import torch

from oracle import ltu_oracle as O
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


def _reference(qkv: torch.Tensor, dout: torch.Tensor, heads: int):
    """fp64 autograd: returns (dqkv [B,N,3C], ctx [B,h,32,32])."""
    B, N, C3 = qkv.shape
    C = C3 // 3
    x = qkv.double().clone().requires_grad_(True)
    q, k, v = (x[..., i * C:(i + 1) * C].view(B, N, heads, 32).transpose(1, 2) for i in range(3))
    out = O.efficient_attention(q, k, v).transpose(1, 2).reshape(B, N, C)
    out.backward(dout.double())
    with torch.no_grad():
        ctx = torch.softmax(k, dim=-2).transpose(-1, -2) @ v
    return x.grad, ctx


CASES = [(2, 50, 4, 1.0), (1, 777, 8, 1.0), (3, 4320, 8, 1.0), (2, 64, 1, 1.0), (1, 1000, 2, 1.0), (2, 2048, 4, 6.0),
         (1, 33, 8, 1.0)]


@pytest.mark.parametrize("B,N,heads,scale", CASES)
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 1e-2)])
def test_attention_backward_matches_autograd(B, N, heads, scale, dtype, tol):
    from lintransunet_b200 import ops
    C = heads * 32
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + N + heads)
    qkv = (torch.randn(B, N, 3 * C, device="cuda", generator=g) * scale).to(dtype)
    dout = torch.randn(B, N, C, device="cuda", generator=g).to(dtype)
    want, ctx = _reference(qkv, dout, heads)
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    got = ops.linear_attention_bwd(q, k, v, ctx.float().contiguous(), dout, heads)
    assert got.shape == (B, N, 3 * C) and got.dtype == dtype
    errs = [rel_err(got[..., i * C:(i + 1) * C], want[..., i * C:(i + 1) * C]) for i in range(3)]
    print(f"\n[attn bwd {dtype} B={B} N={N} h={heads} x{scale}] rel err dq {errs[0]:.2e} dk {errs[1]:.2e} dv {errs[2]:.2e}")
    assert max(errs) <= tol
    again = ops.linear_attention_bwd(q, k, v, ctx.float().contiguous(), dout, heads)
    assert torch.equal(got, again)                                           # fixed-order merges


def test_attention_backward_with_the_forward_kernels_ctx():
    """ctx straight from ops.kv_reduce (what a training step would pass), fp32 path."""
    from lintransunet_b200 import ops
    B, N, heads = 2, 3000, 4
    C = heads * 32
    g = torch.Generator(device="cuda").manual_seed(7)
    qkv = torch.randn(B, N, 3 * C, device="cuda", generator=g)
    dout = torch.randn(B, N, C, device="cuda", generator=g)
    want, _ = _reference(qkv, dout, heads)
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    got = ops.linear_attention_bwd(q, k, v, ops.kv_reduce(k, v, heads), dout, heads)
    assert rel_err(got, want) <= 1e-4
