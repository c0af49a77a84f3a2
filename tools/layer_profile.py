#!/usr/bin/env python
"""Per-launch device times of ONE eager bf16 forward (batch 8 of 128^3, 3 classes) in launch order, with the algorithmic
GB/s and TFLOP/s of every launch (ops.KernelProfiler: CUDA events around each native launch; the GPU is parked on a spin
kernel first so that the events bracket device time, not launch latency)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lintransunet_b200 import MaskTransUnet, ops  # noqa: E402

CFG = dict(num_layers=[16, 32, 64, 128, 256], roi_size_list=[100, 65, 40, 25, 10],
           is_roi_list=[False, True, True, True, True], dim_input=1, dim_output=3)


def main():
    torch.manual_seed(0)
    m = MaskTransUnet(**CFG).cuda().eval()
    m.use_cuda_graphs = False
    x = torch.randn(8, 1, 128, 128, 128, device="cuda")
    with torch.autocast("cuda", dtype=torch.bfloat16):
        for _ in range(2):
            m.predict_labels(x)
        torch.cuda.synchronize()
        prof = ops.KernelProfiler()
        ops.set_profiler(prof)
        torch.cuda._sleep(int(40e-3 * 1.9e9))
        m.predict_labels(x)
        torch.cuda.synchronize()
        ops.set_profiler(None)
    print("| # | kernel | us | GB/s (algorithmic) | TFLOP/s (algorithmic) | TFLOP/s (executed) |")
    print("|---:|---|---:|---:|---:|---:|")
    tot = {}
    for i, (name, e0, e1, nbytes, flops, fexec) in enumerate(prof.records):
        ms = e0.elapsed_time(e1)
        tot[name] = tot.get(name, 0.0) + ms
        if name.startswith("conv") or ms > 0.05:
            print(f"| {i} | {name} | {ms * 1e3:.1f} | {nbytes / ms / 1e6:.0f} | {flops / ms / 1e9:.1f} | {fexec / ms / 1e9:.1f} |")
    print("\n| kernel | ms per forward |\n|---|---:|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"| {k} | {v:.3f} |")
    print(f"| all profiled | {sum(tot.values()):.3f} |")


if __name__ == "__main__":
    main()
