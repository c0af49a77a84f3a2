"""Device times of the row-wise / element-wise backward kernels at the model's shapes (batch 8 of 128^3 patches), bf16,
CUDA events, inputs rotated beyond the L2 size.

    python tools/bwd_probe.py > gpurun_out/bwd_probe.md"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lintransunet_b200 import ops  # noqa: E402


def timeit(fn, nbuf, reps=12):
    for i in range(3):
        fn(i % nbuf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i % nbuf)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def main():
    peak = 6536.4
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    dt = torch.bfloat16
    print(f"| kernel | shape | us | algorithmic MB | GB/s | of the copy peak ({peak:.0f} GB/s) |")
    print("|---|---|---:|---:|---:|---:|")

    def row(name, shape, us, nbytes):
        print(f"| {name} | {shape} | {us:.1f} | {nbytes / 1e6:.0f} | {nbytes / us / 1e3:.0f} | {nbytes / us / 1e3 / peak:.1%} |", flush=True)

    for rows, C in ((8 * 57408, 128), (8 * 10752, 256)):
        nbuf = max(2, (300 << 20) // (rows * C * 2 * 4) + 1)
        bufs = [tuple(torch.randn(rows, C, device="cuda").to(dt) for _ in range(3)) for _ in range(nbuf)]
        gamma = torch.ones(C, device="cuda")
        us = timeit(lambda i: ops.add_layernorm_bwd(bufs[i][0], bufs[i][1], bufs[i][2], gamma, 1e-6), nbuf)
        row("add_layernorm_bwd (+ finalize)", f"{rows} x {C}", us, 4 * rows * C * 2)
        del bufs
    for rows, C in ((8 * 57408, 256), (8 * 10752, 512)):
        nbuf = max(2, (300 << 20) // (rows * C * 2 * 3) + 1)
        bufs = [tuple(torch.randn(rows, C, device="cuda").to(dt) for _ in range(2)) for _ in range(nbuf)]
        us = timeit(lambda i: ops.gelu_bwd(bufs[i][0], bufs[i][1]), nbuf)
        row("gelu_bwd", f"{rows} x {C}", us, 3 * rows * C * 2)
        del bufs
    for (B, H, W, D, C) in ((8, 64, 64, 128, 16), (8, 32, 32, 128, 32), (8, 16, 16, 64, 64)):
        V = H * W * D
        nbuf = max(2, (300 << 20) // (B * V * C * 2 * 3) + 1)
        bufs = [tuple(torch.randn(B, H, W, D, C, device="cuda").to(dt) for _ in range(2)) for _ in range(nbuf)]
        stats = ops.chan_stats(bufs[0][0])
        us = timeit(lambda i: ops.instnorm_bwd(bufs[i][0], stats, bufs[i][1]), nbuf)
        row("instnorm_bwd (partials + finalize + apply)", f"{B} x {H}x{W}x{D} x {C}", us, 5 * B * V * C * 2)
        del bufs
    B, H, W, D, C = 8, 39, 23, 64, 128
    bufs = [tuple(torch.randn(B, H, W, D, C, device="cuda").to(dt) for _ in range(2)) for _ in range(4)]
    w27 = torch.randn(27, C, device="cuda")
    us = timeit(lambda i: ops.posenc_dwconv3_bwd(bufs[i][0], bufs[i][1], w27), 4)
    row("posenc_dwconv3_bwd (dx + wgrad + finalize)", f"{B} x {H}x{W}x{D} x {C}", us, 5 * B * H * W * D * C * 2)


if __name__ == "__main__":
    main()
