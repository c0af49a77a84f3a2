#!/usr/bin/env python
"""Linear layers of one d_model-256 encoder layer (and bridge 1's K/V projection) at the model's token counts, batch 8 of
128^3: ltu_linear_fused (TMA + tcgen05, fused epilogues) against cuBLAS (F.linear) + the separate gelu / add_layernorm
kernels it replaces.  CUDA-graph replay over rotating buffers larger than L2, CUDA events."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lintransunet_b200 import ops  # noqa: E402
from tools.attn_core_probe import graph_time  # noqa: E402

PEAK_BW, PEAK_TF = 6536.4, 1387.9
bf = torch.bfloat16


def main():
    print("| layer | rows | K | N | fused us | GB/s | of HBM | TFLOP/s | cuBLAS(+kernels) us | speed-up |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    tot_f = tot_c = 0.0
    quick = os.environ.get("LTU_PROBE_QUICK") == "1"          # ablation runs (LTU_LIN_MODE): bridge 2 only, no cuBLAS leg
    shapes = ((8 * 10752, "b2"),) if quick else ((8 * 10752, "b2"), (8 * 4320, "b3"), (8 * 512, "b4"), (8 * 57408, "b1"))
    if quick:
        print(f"# LTU_LIN_MODE={os.environ.get('LTU_LIN_MODE', '0')}")
    for rows, tag in shapes:
        cases = [("kv", 128, 256, 0)] if tag == "b1" else [("qkv", 256, 768, 0), ("o+ln", 256, 256, 2), ("ffn1+gelu", 256, 512, 1), ("ffn1 (bias only, for comparison)", 256, 512, 0),
                                                           ("ffn2+ln", 512, 256, 2)]
        for name, K, N, epi in cases:
            io = rows * (K + N) * 2 + (rows * N * 2 * 3 if epi == 2 else 0)
            nbuf = max(2, min(16, int(500e6 // io) + 1))
            xs = [torch.randn(rows, K, device="cuda").to(bf) for _ in range(nbuf)]
            rh = [torch.randn(rows, N, device="cuda").to(bf) for _ in range(nbuf)] if epi == 2 else None
            rl = [(torch.randn(rows, N, device="cuda") * 2 ** -9).to(bf) for _ in range(nbuf)] if epi == 2 else None
            w = (torch.randn(N, K, device="cuda") * 0.05).to(bf)
            b32 = torch.randn(N, device="cuda")
            b16 = b32.to(bf)
            g, be = torch.ones(N, device="cuda"), torch.zeros(N, device="cuda")

            def fused(i):
                if epi == 2:
                    return ops.linear_fused(xs[i], w, b32, 2, rh[i], rl[i], g, be)
                return ops.linear_fused(xs[i], w, b32, epi)

            def cublas(i):
                y = F.linear(xs[i], w, b16)
                if epi == 1:
                    return ops.gelu_(y)
                if epi == 2:
                    return ops.add_layernorm_split(rh[i], rl[i], y, g, be)
                return y

            tf_ = graph_time(fused, nbuf)
            tc_ = tf_ if quick else graph_time(cublas, nbuf)
            if tag != "b1" and "comparison" not in name:
                tot_f += tf_ * 8; tot_c += tc_ * 8
            print(f"| {tag} {name} | {rows} | {K} | {N} | {tf_:.1f} | {io / tf_ / 1e3:.0f} | {io / tf_ / 1e3 / PEAK_BW * 100:.1f}% | "
                  f"{2 * rows * K * N / tf_ / 1e6:.0f} | {tc_:.1f} | {tc_ / tf_:.2f}x |", flush=True)
    print(f"\nper forward (8 layers x bridges 2-4, B=8): fused {tot_f:.0f} us, cuBLAS + separate kernels {tot_c:.0f} us")
    # bridge 1: K/V projection + kv_reduce as one launch (K, V never written) against the two separate native launches
    B, N, C, h = 8, 57408, 128, 4
    xs = [torch.randn(B, N, C, device="cuda").to(bf) for _ in range(4)]
    w = (torch.randn(2 * C, C, device="cuda") * 0.05).to(bf)
    b32 = torch.randn(2 * C, device="cuda")
    t_f = graph_time(lambda i: ops.kv_project_reduce(xs[i], w, b32, h), 4)

    def sep(i):
        kv = ops.linear_fused(xs[i], w, b32)
        return ops.kv_reduce(kv[..., :C], kv[..., C:], h)
    t_s = graph_time(sep, 4)
    print(f"\nbridge 1 K/V half (B=8, N=57408, C=128): kv_project_reduce {t_f:.1f} us (x read once: "
          f"{B * N * C * 2 / t_f / 1e3:.0f} GB/s) vs linear_fused + kv_reduce {t_s:.1f} us ({t_s / t_f:.2f}x)")


if __name__ == "__main__":
    main()
