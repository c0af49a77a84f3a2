"""Time the fused attention-output kernel against the separate path on one B200 (CUDA events)."""
import sys
import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from lintransunet_b200 import ops  # noqa: E402
from tools.ffn_probe import timeit  # noqa: E402


def main():
    C, h = 128, 4
    torch.manual_seed(0)
    for (B, N) in ((8, 57408), (1, 57408)):
        x = torch.randn(B, N, C, device="cuda").to(torch.bfloat16)
        wqkv = (torch.randn(3 * C, C, device="cuda") * 0.1).to(torch.bfloat16)
        bqkv = (torch.randn(3 * C, device="cuda") * 0.1).to(torch.bfloat16)
        wo = (torch.randn(C, C, device="cuda") * 0.1).to(torch.bfloat16)
        bo = (torch.randn(C, device="cuda") * 0.1).to(torch.bfloat16)
        g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
        wq, wkv = wqkv[:C].contiguous(), wqkv[C:].contiguous()
        bq, bkv = bqkv[:C].float().contiguous(), bqkv[C:].contiguous()

        def separate():
            qkv = F.linear(x, wqkv, bqkv)
            q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
            ctx = ops.kv_reduce(k, v, h)
            att = ops.q_readout(q, ctx, h)
            o = F.linear(att, wo, bo)
            return ops.add_layernorm(x, o, g, b, 1e-6)

        def fused():
            kv = F.linear(x, wkv, bkv)
            ctx = ops.kv_reduce(kv[..., :C], kv[..., C:], h)
            return ops.attn_out_fused(x, wq, bq, ops.ctx_pack(ctx), wo, bo.float(), g, b, h)

        kv = F.linear(x, wkv, bkv)
        ctx16 = ops.ctx_pack(ops.kv_reduce(kv[..., :C], kv[..., C:], h))
        only = lambda: ops.attn_out_fused(x, wq, bq, ctx16, wo, bo.float(), g, b, h)
        ya, yb = separate(), fused()
        err = (ya.float() - yb.float()).abs().max().item()
        ts, tf, to = timeit(separate), timeit(fused), timeit(only)
        print(f"B={B} N={N}: separate {ts:.1f} us, fused path {tf:.1f} us, attn_out_fused alone {to:.1f} us "
              f"({2 * B * N * C * 2 / to / 1e3:.0f} GB/s of x+y), max|diff| {err:.4f}", flush=True)


if __name__ == "__main__":
    main()
