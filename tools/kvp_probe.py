"""kv_project_reduce at the benchmark's bridge-1 shape (B=8, N=57408, C=128): a few calls, for `ncu --metrics gpu__time_duration.sum`."""
import sys
import torch
sys.path.insert(0, ".")
from lintransunet_b200 import ops  # noqa: E402

torch.manual_seed(0)
B, N, C, h = 8, 57408, 128, 4
x = torch.randn(B, N, C, device="cuda").to(torch.bfloat16)
w = (torch.randn(2 * C, C, device="cuda") * 0.1).to(torch.bfloat16)
b = torch.randn(2 * C, device="cuda") * 0.1
wo = (torch.randn(C, C, device="cuda") * 0.1).to(torch.bfloat16)
for _ in range(4):
    ctx, wb = ops.kv_project_reduce(x, w, b, h, w_o=wo)
torch.cuda.synchronize()
print("ok", float(ctx.abs().mean()))
