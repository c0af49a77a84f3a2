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
    torch.manual_seed(0)
    for C, rows in ((128, 8 * 57408), (128, 57408), (256, 8 * 10752), (256, 8 * 4320), (256, 8 * 512)):
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
        print(f"C={C} rows={rows}: separate {ts:.1f} us, fused {tf:.1f} us ({gbs:.0f} GB/s of x+y, "
              f"{2 * rows * C * 2 * C * 2 / tf / 1e6:.0f} TFLOP/s), max|diff| {err:.4f}", flush=True)




def trace():
    """Pipeline trace of CTA 0 (clock64 deltas, cycles)."""
    from ctypes import c_void_p
    from lintransunet_b200 import _native
    C, rows = 128, 8 * 57408
    x = torch.randn(rows, C, device="cuda").to(torch.bfloat16)
    w1 = (torch.randn(2 * C, C, device="cuda") * 0.1).to(torch.bfloat16)
    w2 = (torch.randn(C, 2 * C, device="cuda") * 0.1).to(torch.bfloat16)
    b1 = torch.randn(2 * C, device="cuda") * 0.1
    b2 = torch.randn(C, device="cuda") * 0.1
    g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    y = torch.empty_like(x)
    tr = torch.zeros(2, 64, 8, dtype=torch.int64, device="cuda")
    P = lambda t: c_void_p(t.data_ptr())
    for _ in range(2):
        tr.zero_()
        rc = _native.lib().ltu_ffn_fused_trace(P(x), rows, C, P(w1), P(b1), P(w2), P(b2), P(g), P(b), 1e-6, P(y), P(tr),
                                               c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0
    torch.cuda.synchronize()
    t = tr.cpu()
    t0 = int(t[t > 0].min())
    for tile in range(2, 10):
        a = [int(v) - t0 if v else -1 for v in t[0, tile, :4]]
        c = [int(v) - t0 if v else -1 for v in t[1, tile, :7]]
        print(f"tile {tile}: e1 wait {a[0]} ready {a[1]} gelu_done {a[2]} g2_issue {a[3]} | e2 wait {c[0]} acc2 {c[1]} drained {c[2]} xland {c[3]} stats {c[4]} xchg {c[5]} stored {c[6]}")


def ablate():
    from ctypes import c_void_p
    from lintransunet_b200 import _native
    C, rows = 128, 8 * 57408
    x = torch.randn(rows, C, device="cuda").to(torch.bfloat16)
    w1 = (torch.randn(2 * C, C, device="cuda") * 0.1).to(torch.bfloat16)
    w2 = (torch.randn(C, 2 * C, device="cuda") * 0.1).to(torch.bfloat16)
    b1 = torch.randn(2 * C, device="cuda") * 0.1
    b2 = torch.randn(C, device="cuda") * 0.1
    g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    y = torch.empty_like(x)
    P = lambda t: c_void_p(t.data_ptr())
    for mode in (0, 3, 7, 11, 15, 19, 31):
        fn = lambda: _native.lib().ltu_ffn_fused_trace(P(x), rows, C, P(w1), P(b1), P(w2), P(b2), P(g), P(b), 1e-6, P(y),
                                                       c_void_p(mode), c_void_p(torch.cuda.current_stream().cuda_stream))
        print(f"mode {mode}: {timeit(fn):.1f} us", flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "trace":
    trace()
elif __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "ablate":
    ablate()
elif __name__ == "__main__":
    main()
