"""Which shared-memory HALO addressing does a SWIZZLE_128B UMMA operand descriptor accept on B200?
Prints, per (start row offset, rows between 8-row groups, base_offset field), whether A . B^T is exact."""
import ctypes
import os
import subprocess
import sys
from ctypes import c_void_p

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lintransunet_b200 import build as _b  # noqa: E402


def probe_lib():
    """The probe is a development tool, NOT part of libltu_b200.so: it is compiled here into tools/libltu_probe.so
    and linked against the product library for the helpers it needs (error string, tensor-map encoder)."""
    so = os.path.join(ROOT, "tools", "libltu_probe.so")
    src = os.path.join(ROOT, "tools", "csrc", "umma_probe.cu")
    lib = _b.build()                                  # helpers (error string, tensor-map encoder) come from the product .so
    if not os.path.exists(so) or max(os.path.getmtime(src), os.path.getmtime(lib)) > os.path.getmtime(so):
        subprocess.check_call([_b.NVCC, *_b.FLAGS, "-I", _b.CSRC, "-shared", "-o", so, src, "-L", os.path.dirname(lib),
                               "-lltu_b200", "-Xlinker", "-rpath," + os.path.dirname(lib)])
    h = ctypes.CDLL(so)
    h.ltu_debug_umma_probe.restype = ctypes.c_int
    h.ltu_debug_umma_probe.argtypes = [c_void_p, ctypes.c_int, c_void_p, c_void_p] + [ctypes.c_int] * 4 + [c_void_p]
    return h


LIB = probe_lib()

torch.manual_seed(0)
R = 256
out = torch.empty(128, 64, device="cuda")
P = lambda t: c_void_p(t.data_ptr())
for CW in (64, 32, 16):
  print(f"--- {CW} channels per row ({2 * CW}-byte rows, SWIZZLE_{2 * CW}B)")
  g = torch.randint(-4, 5, (R, CW), device="cuda").to(torch.bfloat16)
  w = torch.randint(-2, 3, (64, CW), device="cuda").to(torch.bfloat16)
  for off, sbo in ((0, 8), (8, 8), (1, 8), (3, 8), (0, 10), (2, 10), (5, 10), (11, 10), (0, 16), (7, 12)):
   for _ in (0,):
    rows = torch.tensor([off + (r // 8) * sbo + r % 8 for r in range(128)], device="cuda")
    ref = g[rows].float() @ w.float().t()
    res = []
    for ubo in (0, 1):
        out.zero_()
        rc = LIB.ltu_debug_umma_probe(P(g), R, P(w), P(out), off, sbo, ubo, CW, c_void_p(torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        bad = (out != ref).any(dim=1)
        res.append("exact" if rc == 0 and not bad.any() else f"rc={rc} wrong rows {int(bad.sum())}/128 (first {int(bad.nonzero()[0]) if bad.any() else -1})")
    print(f"start row {off:2d}, group stride {sbo:2d} rows: base_offset=0 -> {res[0]}; base_offset=(addr>>7)&7 -> {res[1]}", flush=True)
