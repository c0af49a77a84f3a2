#!/usr/bin/env python
"""Quick device timings (CUDA events) of the forward and of the attention-core kernels."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lintransunet_b200 import MaskTransUnet, ops, _native  # noqa: E402


def timeit(fn, warm=3, it=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


def main():
    torch.manual_seed(0)
    m = MaskTransUnet([16, 32, 64, 128, 256], [100, 65, 40, 25, 10], [False, True, True, True, True], 1, 3).cuda().eval()
    for prec, B, S in (("bf16", 1, 128), ("bf16", 8, 128), ("fp32", 1, 128), ("bf16", 2, 96)):
        m.precision = prec
        x = torch.randn(B, 1, S, S, S, device="cuda")
        n0 = _native.launch_count()
        m(x)
        n1 = _native.launch_count()
        t0 = time.perf_counter()
        ms = timeit(lambda: m.predict_labels(x), warm=2, it=5)
        wall = (time.perf_counter() - t0) / 7 * 1e3
        print(f"forward {prec} B={B} {S}^3: {ms:.2f} ms/iter (wall {wall:.2f}) -> {B*S**3/ms*1e3:.3e} voxels/s; "
              f"{n1-n0} native launches/forward", flush=True)
    for fused in (True, False):
        m.precision, m.use_fused_linear = "bf16", fused
        x = torch.randn(8, 1, 128, 128, 128, device="cuda")
        ms = timeit(lambda: m.predict_labels(x), warm=2, it=5)
        print(f"forward bf16 B=8 128^3 fused_linear={fused}: {ms:.2f} ms/iter", flush=True)
    m.use_fused_linear = False
    import torch.nn.functional as F
    for (M, C) in ((8 * 57408, 128), (8 * 10752, 256)):
        t = torch.randn(M, C, device="cuda").to(torch.bfloat16)
        lin1, lin2 = torch.nn.Linear(C, 2 * C).cuda(), torch.nn.Linear(2 * C, C).cuda()
        w1, b1 = lin1.weight.detach().to(torch.bfloat16), lin1.bias.detach().to(torch.bfloat16)
        w2, b2 = lin2.weight.detach().to(torch.bfloat16), lin2.bias.detach().to(torch.bfloat16)
        p1, p2 = ops.pack_linear_tc(lin1.weight), ops.pack_linear_tc(lin2.weight)
        f1, f2 = lin1.bias.detach().float(), lin2.bias.detach().float()
        g, be = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")

        def unfused():
            f = ops.gelu_(F.linear(t, w1, b1))
            return ops.add_layernorm(t, F.linear(f, w2, b2), g, be)

        def fused():
            f = ops.linear_tc(t, p1, f1, 2 * C, ops.EPI_GELU)
            return ops.linear_tc(f, p2, f2, C, ops.EPI_RES_LN, residual=t, gamma=g, beta=be)

        print(f"FFN M={M} C={C}: cuBLAS+gelu+LN {timeit(unfused)*1e3:.1f} us, fused tcgen05 {timeit(fused)*1e3:.1f} us", flush=True)
    # attention core at the model's token counts
    for dt in (torch.bfloat16, torch.float32):
        for (B, N, h) in ((8, 57408, 4), (8, 10752, 8), (8, 4320, 8), (8, 512, 8), (1, 57408, 4), (1, 32768, 8)):
            C = 32 * h
            qkv = torch.randn(B, N, 3 * C, device="cuda").to(dt)
            q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
            ctx = ops.kv_reduce(k, v, h)
            t_kv = timeit(lambda: ops.kv_reduce(k, v, h))
            t_q = timeit(lambda: ops.q_readout(q, ctx, h))
            es = qkv.element_size()
            by_kv, by_q = 2 * B * N * C * es, 2 * B * N * C * es
            print(f"attn {str(dt)[6:]} B={B} N={N} h={h}: kv_reduce {t_kv*1e3:.1f} us ({by_kv/t_kv/1e6:.0f} GB/s) "
                  f"q_readout {t_q*1e3:.1f} us ({by_q/t_q/1e6:.0f} GB/s) core {(by_kv+by_q)/(t_kv+t_q)/1e6:.0f} GB/s",
                  flush=True)
    # per-stage profile of one bf16 B=8 forward with the torch profiler
    from torch.profiler import profile, ProfilerActivity
    m.precision = "bf16"
    x = torch.randn(8, 1, 128, 128, 128, device="cuda")
    m.predict_labels(x)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        m.predict_labels(x)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))


if __name__ == "__main__":
    main()
