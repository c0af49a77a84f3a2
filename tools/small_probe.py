"""Time the small latency-bound kernels (posenc, kv_reduce+combine, instnorm_finalize) on one B200."""
import sys
import torch

sys.path.insert(0, ".")
from lintransunet_b200 import ops  # noqa: E402
from tools.ffn_probe import timeit  # noqa: E402

torch.manual_seed(0)
for shape in ((8, 39, 23, 64, 128), (8, 24, 14, 32, 256), (8, 15, 9, 32, 256), (8, 4, 4, 32, 256)):
    x = torch.randn(*shape, device="cuda").to(torch.bfloat16)
    w = torch.randn(27, shape[-1], device="cuda") * 0.1
    b = torch.randn(shape[-1], device="cuda") * 0.1
    t = timeit(lambda: ops.posenc_dwconv3(x, w, b))
    print(f"posenc {shape}: {t:.1f} us ({2 * x.numel() * 2 / t / 1e3:.0f} GB/s)", flush=True)
for (B, N, h) in ((8, 57408, 4), (8, 10752, 8), (8, 4320, 8), (8, 512, 8)):
    C = 32 * h
    qkv = torch.randn(B, N, 3 * C, device="cuda").to(torch.bfloat16)
    k, v = qkv[..., C:2 * C], qkv[..., 2 * C:]
    t = timeit(lambda: ops.kv_reduce(k, v, h))
    print(f"kv_reduce+combine B={B} N={N} h={h}: {t:.1f} us ({2 * B * N * C * 2 / t / 1e3:.0f} GB/s)", flush=True)
for (B, tiles, C) in ((8, 4096, 16), (8, 1024, 32), (8, 256, 64)):
    part = torch.randn(B, tiles, C, 2, device="cuda")
    t = timeit(lambda: ops.instnorm_finalize(part, tiles * 128))
    print(f"instnorm_finalize B={B} tiles={tiles} C={C}: {t:.1f} us", flush=True)
