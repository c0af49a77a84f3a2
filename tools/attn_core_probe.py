#!/usr/bin/env python
"""kv_reduce / q_readout device time at the model's token counts (batch 8 of 128^3), CUDA-graph replay over rotating
buffers larger than L2, CUDA events.  Run twice: LTU_ATTN_STREAM=1 (TMA streaming kernels) and =0 (cp.async kernels)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lintransunet_b200 import ops  # noqa: E402

PEAK = 6536.4


def graph_time(fn, nbuf, reps=3):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for i in range(nbuf):
            fn(i)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(nbuf):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * nbuf) * 1e3          # us per call


def main():
    mode = os.environ.get("LTU_ATTN_STREAM", "1")
    print(f"# LTU_ATTN_STREAM={mode}")
    print("| B | heads | N | kv_reduce us | GB/s | of peak | q_readout us | GB/s | of peak | core us |")
    print("|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    tot = {"kv": 0.0, "q": 0.0, "bytes_kv": 0, "bytes_q": 0}
    for B, h, N in ((8, 4, 57408), (8, 8, 10752), (8, 8, 4320), (8, 8, 512), (7, 8, 10752), (1, 8, 10752), (2, 4, 43056)):
        C = 32 * h
        nbytes = B * N * 3 * C * 2
        nbuf = max(2, min(24, int(400e6 // nbytes) + 1))
        bufs = [torch.randn(B, N, 3 * C, device="cuda").to(torch.bfloat16) for _ in range(nbuf)]
        ctxs = [ops.kv_reduce(b[..., C:2 * C], b[..., 2 * C:], h) for b in bufs]
        t_kv = graph_time(lambda i: ops.kv_reduce(bufs[i][..., C:2 * C], bufs[i][..., 2 * C:], h), nbuf)
        t_q = graph_time(lambda i: ops.q_readout(bufs[i][..., :C], ctxs[i], h), nbuf)
        t_core = graph_time(lambda i: ops.q_readout(bufs[i][..., :C], ops.kv_reduce(bufs[i][..., C:2 * C], bufs[i][..., 2 * C:], h), h), nbuf)
        by = 2 * B * N * C * 2
        print(f"| {B} | {h} | {N} | {t_kv:.1f} | {by / t_kv / 1e3:.0f} | {by / t_kv / 1e3 / PEAK * 100:.1f}% | {t_q:.1f} | "
              f"{by / t_q / 1e3:.0f} | {by / t_q / 1e3 / PEAK * 100:.1f}% | {t_core:.1f} |", flush=True)
        if B == 8:
            tot["kv"] += t_kv * 8; tot["bytes_kv"] += by * 8
            if h == 8:
                tot["q"] += t_q * 8; tot["bytes_q"] += by * 8
    print(f"\nper forward (8 layers per bridge, B=8): kv_reduce {tot['kv']:.0f} us = {tot['bytes_kv'] / tot['kv'] / 1e3:.0f} GB/s "
          f"({tot['bytes_kv'] / tot['kv'] / 1e3 / PEAK * 100:.1f}% of {PEAK:.0f}); q_readout (bridges 2-4) {tot['q']:.0f} us = "
          f"{tot['bytes_q'] / tot['q'] / 1e3:.0f} GB/s ({tot['bytes_q'] / tot['q'] / 1e3 / PEAK * 100:.1f}%)")


if __name__ == "__main__":
    main()
