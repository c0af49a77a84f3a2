"""MN-major UMMA operands on B200: which (LBO, SBO, K-step advance) reads a [tokens x channels] SWIZZLE_128B tile with the
TOKEN axis as the K dimension?  out = P^T . X for P, X bf16 [128 tokens][128 channels] (two 64-channel TMA blocks each)."""
import ctypes
import os
import sys
from ctypes import c_void_p

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv = sys.argv[:1]
import tools.umma_probe as base  # noqa: E402  (builds tools/libltu_probe.so; its own sweep prints first)

LIB = base.LIB
LIB.ltu_debug_umma_probe_mn.restype = ctypes.c_int
LIB.ltu_debug_umma_probe_mn.argtypes = [c_void_p, c_void_p, c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_void_p]
torch.manual_seed(1)
p = torch.randint(-3, 4, (128, 128), device="cuda").to(torch.bfloat16)
x = torch.randint(-3, 4, (128, 128), device="cuda").to(torch.bfloat16)
ref = p.float().t() @ x.float()
out = torch.empty(128, 128, device="cuda")
P = lambda t: c_void_p(t.data_ptr())
for lbo, sbo, kadv in ((16384, 1024, 2048), (1024, 16384, 2048), (16384, 1024, 256), (16384, 2048, 2048), (8192, 1024, 2048)):
    out.zero_()
    rc = LIB.ltu_debug_umma_probe_mn(P(p), P(x), P(out), lbo, sbo, kadv, c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    bad = (out != ref)
    print(f"MN-major A and B: LBO {lbo:5d} SBO {sbo:5d} K-step advance {kadv:4d} B -> "
          f"{'exact' if rc == 0 and not bad.any() else f'rc={rc} wrong {int(bad.sum())}/16384'}", flush=True)
