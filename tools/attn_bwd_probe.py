"""Device time of the attention-core backward (ltu_attn_bwd: reduce + combine + apply) at the model's token counts,
bf16 and fp32, CUDA events, inputs rotated over more than the L2 size.

    python tools/attn_bwd_probe.py > gpurun_out/attn_bwd_probe.md"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lintransunet_b200 import ops  # noqa: E402


def main():
    print("| dtype | B | heads | N | us | algorithmic MB (10 N C E) | GB/s | floor MB (7 N C E) |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|")
    for dtype in (torch.bfloat16, torch.float32):
        for B, heads, N in ((8, 4, 57408), (8, 8, 10752), (8, 8, 4320), (8, 8, 512), (2, 4, 43056)):
            C = heads * 32
            es = 2 if dtype == torch.bfloat16 else 4
            nbuf = max(2, min(16, (300 << 20) // (4 * B * N * C * es) + 1))
            bufs = [(torch.randn(B, N, 3 * C, device="cuda").to(dtype), torch.randn(B, N, C, device="cuda").to(dtype))
                    for _ in range(nbuf)]
            ctx = ops.kv_reduce(bufs[0][0][..., C:2 * C], bufs[0][0][..., 2 * C:], heads)

            def run(i):
                qkv, g = bufs[i % nbuf]
                return ops.linear_attention_bwd(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], ctx, g, heads)
            for i in range(3):
                run(i)
            torch.cuda.synchronize()
            reps = 10
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(reps):
                run(i)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / reps
            nb = 10 * B * N * C * es
            print(f"| {str(dtype)[6:]} | {B} | {heads} | {N} | {us:.1f} | {nb / 1e6:.0f} | {nb / us / 1e3:.0f} | {0.7 * nb / 1e6:.0f} |",
                  flush=True)
            del bufs
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
