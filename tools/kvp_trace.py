"""Pipeline trace of CTA 0 of kv_project2_kernel at the benchmark's bridge-1 shape (clock64 stamps, cycles relative to the
first stamp): where does a tile's time go?   python tools/kvp_trace.py"""
import os
import sys
import torch
sys.path.insert(0, ".")
trace = torch.zeros(3 * 64 * 8, dtype=torch.int64, device="cuda")
os.environ["LTU_KVP_TRACE_PTR"] = str(trace.data_ptr())
from lintransunet_b200 import ops  # noqa: E402

torch.manual_seed(0)
B, N, C, h = 8, 57408, 128, 4
x = torch.randn(B, N, C, device="cuda").to(torch.bfloat16)
w = (torch.randn(2 * C, C, device="cuda") * 0.1).to(torch.bfloat16)
b = torch.randn(2 * C, device="cuda") * 0.1
for _ in range(3):
    trace.zero_()
    ops.kv_project_reduce(x, w, b, h)
torch.cuda.synchronize()
t = trace.cpu().reshape(3, 64, 8)
t0 = int(t[t > 0].min())
print("kernel entry", int(t[0, 63, 7]) - t0, "all roles done", int(t[1, 63, 7]) - t0, "cycles (SM clock)")
rel = lambda v: int(v) - t0 if int(v) > 0 else -1
print("tile | TMA issue | Kacc free, x landed (MMA1) | wait P, P ready (MMA2) || softmax warp: wait Kacc, Kacc ready, read+check, vote, P buf free, P published")
for i in range(26):
    print(f"{i:3d} | {rel(t[0, i, 0]):7d} | {rel(t[1, i, 0]):7d} {rel(t[1, i, 1]):7d} | {rel(t[1, i, 2]):7d} {rel(t[1, i, 3]):7d} || "
          + " ".join(f"{rel(t[2, i, k]):7d}" for k in range(7)))
