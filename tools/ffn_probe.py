"""Time the fused FFN kernel against cuBLAS + gelu + add_layernorm on one B200 (CUDA events)."""
import sys
import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from lintransunet_b200 import ops  # noqa: E402


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    C = 128
    torch.manual_seed(0)
    for rows in (8 * 57408, 57408, 8 * 4320):
        x = torch.randn(rows, C, device="cuda").to(torch.bfloat16)
        w1 = (torch.randn(2 * C, C, device="cuda") * 0.1).to(torch.bfloat16)
        w2 = (torch.randn(C, 2 * C, device="cuda") * 0.1).to(torch.bfloat16)
        b1 = torch.randn(2 * C, device="cuda") * 0.1
        b2 = torch.randn(C, device="cuda") * 0.1
        g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
        b1h, b2h = b1.to(torch.bfloat16), b2.to(torch.bfloat16)

        def separate():
            f = ops.gelu_(F.linear(x, w1, b1h))
            f = F.linear(f, w2, b2h)
            return ops.add_layernorm(x, f, g, b, 1e-6)

        def fused():
            return ops.ffn_fused(x, w1, b1, w2, b2, g, b, 1e-6)

        ya, yb = separate(), fused()
        err = (ya.float() - yb.float()).abs().max().item()
        ts, tf = timeit(separate), timeit(fused)
        gbs = 2 * rows * C * 2 / tf / 1e3
        print(f"rows={rows}: separate {ts:.1f} us, fused {tf:.1f} us ({gbs:.0f} GB/s of x+y, "
              f"{2 * rows * C * 2 * C * 2 / tf / 1e6:.0f} TFLOP/s), max|diff| {err:.4f}", flush=True)


if __name__ == "__main__":
    main()
