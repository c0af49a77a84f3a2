"""CPU: the algebra the attention-core backward kernels implement (attn_bwd.cu / attn_bwd_tc.cu header) against fp64 autograd
through the oracle's `linear_attention` restatement -- in particular that the column-softmax correction sum_n Ks[n][j] dKs[n][j]
equals sum_e dctx[j][e] ctx[j][e], so K's softmax backward needs no second pass over the tokens."""
import math

import torch

from oracle import ltu_oracle as O


def test_two_pass_backward_formulas_equal_autograd():
    torch.manual_seed(0)
    B, h, N, d = 2, 4, 70, 32
    q, k, v = (torch.randn(B, h, N, d, dtype=torch.float64, requires_grad=True) for _ in range(3))
    out = O.efficient_attention(q, k, v)
    g = torch.randn_like(out)
    out.backward(g)
    with torch.no_grad():
        P = torch.softmax(q, -1)
        Qs = P / math.sqrt(d)
        M = k.max(-2, keepdim=True).values
        S = torch.exp(k - M).sum(-2, keepdim=True)
        Ks = torch.exp(k - M) / S
        ctx = Ks.transpose(-1, -2) @ v
        # pass 1: one reduction over the tokens
        dctx = Qs.transpose(-1, -2) @ g
        t = (dctx * ctx).sum(-1)                                   # [B,h,j]
        # the identity behind the single pass
        dKs = v @ dctx.transpose(-1, -2)
        assert torch.allclose(t, (Ks * dKs).sum(-2), rtol=1e-12, atol=1e-14)
        # pass 2: element-wise per token
        dV = Ks @ dctx
        dK = Ks * (dKs - t.unsqueeze(-2))
        dP = (g @ ctx.transpose(-1, -2)) / math.sqrt(d)
        dQ = P * (dP - (P * dP).sum(-1, keepdim=True))
    for got, ref in ((dQ, q.grad), (dK, k.grad), (dV, v.grad)):
        assert torch.allclose(got, ref, rtol=1e-10, atol=1e-13)
    # K's bias gradient is identically zero: the column softmax is invariant to a per-column shift
    assert float(k.grad.sum(-2).abs().max()) < 1e-12
