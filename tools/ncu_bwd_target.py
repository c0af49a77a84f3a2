"""ncu target: two calls of the attention-core backward at bridge 1's shape (B=8, N=57408, 4 heads, bf16)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lintransunet_b200 import ops  # noqa: E402

B, N, h = 8, 57408, 4
C = h * 32
qkv = torch.randn(B, N, 3 * C, device="cuda").to(torch.bfloat16)
g = torch.randn(B, N, C, device="cuda").to(torch.bfloat16)
ctx = ops.kv_reduce(qkv[..., C:2 * C], qkv[..., 2 * C:], h)
for _ in range(2):
    ops.linear_attention_bwd(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], ctx, g, h)
torch.cuda.synchronize()
