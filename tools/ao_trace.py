"""Pipeline trace of CTA 0 of attn_out128w_kernel at the benchmark's bridge-1 shape (clock64 stamps, cycles relative to the
first stamp).   python tools/ao_trace.py"""
import os
import sys
import torch
sys.path.insert(0, ".")
trace = torch.zeros(4 * 64 * 8, dtype=torch.int64, device="cuda")
os.environ["LTU_AO_TRACE_PTR"] = str(trace.data_ptr())
from lintransunet_b200 import ops  # noqa: E402

torch.manual_seed(0)
B, N, C, h = 8, 57408, 128, 4
x = torch.randn(B, N, C, device="cuda").to(torch.bfloat16)
wq = (torch.randn(C, C, device="cuda") * 0.1).to(torch.bfloat16)
wb = (torch.randn(B, C, C, device="cuda") * 0.1).to(torch.bfloat16)
bq, bo = torch.randn(C, device="cuda") * 0.1, torch.randn(C, device="cuda") * 0.1
g, be = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
for _ in range(3):
    trace.zero_()
    ops.attn_out_fused_w(x, wq, bq, wb, bo, g, be, h)
torch.cuda.synchronize()
t = trace.cpu().reshape(4, 64, 8)
t0 = int(t[t > 0].min())
rel = lambda v: int(v) - t0 if int(v) > 0 else -1
print("tile | TMA: x issued, W_b issued | GQ issued, GO issued || softmax warp: wait Q, Q ready, P published || LN warp: wait GO, GO ready, "
      "stats, partner, row written, store read")
for i in range(26):
    print(f"{i:3d} | {rel(t[0, i, 0]):7d} {rel(t[0, i, 1]):7d} | {rel(t[1, i, 0]):7d} {rel(t[1, i, 1]):7d} || "
          + " ".join(f"{rel(t[2, i, k]):7d}" for k in range(3)) + " || " + " ".join(f"{rel(t[3, i, k]):7d}" for k in range(6)))
