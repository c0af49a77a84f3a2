"""Device time of ltu_conv3d_wgrad on layers of the model at batch 8 of 128^3 patches (bf16), CUDA events.

    python tools/conv_wgrad_probe.py > gpurun_out/conv_wgrad_probe.md"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lintransunet_b200 import ops  # noqa: E402

LAYERS = [  # name, cin, cout, k, stride, input (H, W, D)
    ("enc.block0.conv1", 16, 16, 3, (1, 1, 1), (64, 64, 128)),
    ("enc.block1.conv1", 32, 32, 3, (1, 1, 1), (32, 32, 128)),
    ("enc.block1.conv2", 32, 64, 3, (2, 2, 2), (32, 32, 128)),
    ("enc.block2.conv1", 64, 64, 3, (1, 1, 1), (16, 16, 64)),
    ("enc.block3.conv1", 128, 128, 3, (1, 1, 1), (8, 8, 64)),
    ("b1.down_embed", 32, 128, 3, (2, 2, 2), (78, 46, 128)),
]


def main():
    print("| layer | Cin -> Cout | us | TFLOP/s (2*27*Cin*Cout*B*Vout) |")
    print("|---|---|---:|---:|")
    for name, cin, cout, k, stride, (H, W, D) in LAYERS:
        B = 8
        x = torch.randn(B, H, W, D, cin, device="cuda").to(torch.bfloat16)
        Ho, Wo, Do = ((n + 2 * (k // 2) - k) // s + 1 for n, s in zip((H, W, D), stride))
        dy = torch.randn(B, Ho, Wo, Do, cout, device="cuda").to(torch.bfloat16)
        for _ in range(2):
            ops.conv3d_wgrad(x, dy, k, stride, k // 2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.conv3d_wgrad(x, dy, k, stride, k // 2)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 5
        flops = 2 * k ** 3 * cin * cout * B * Ho * Wo * Do
        print(f"| {name} | {cin} -> {cout} | {us:.0f} | {flops / us / 1e6:.0f} |", flush=True)


if __name__ == "__main__":
    main()
