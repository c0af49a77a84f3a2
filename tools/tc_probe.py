#!/usr/bin/env python
"""Diagnostics for the tcgen05 convolution: identity-weight and random cases vs F.conv3d,
with enough printed detail to tell layout / swizzle / descriptor errors apart."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lintransunet_b200 import ops  # noqa: E402
from lintransunet_b200.unet import _ConvW  # noqa: E402


def run(cin, cin1, cout, stride, up2, shape, mode, B=1, k=3, out_f32=False):
    H, W, D = shape
    conv = torch.nn.Conv3d(cin + cin1, cout, k, stride=stride, padding=k // 2)
    with torch.no_grad():
        if mode == "identity":
            conv.weight.zero_()
            conv.bias.zero_()
            for n in range(min(cout, cin + cin1)):
                conv.weight[n, n, 1, 1, 1] = 1.0
        else:
            conv.weight.copy_(conv.weight.to(torch.bfloat16).float())
    cw = _ConvW(conv, want_tc=True, fold_up2=up2)
    g = torch.Generator().manual_seed(1)
    x0 = torch.randn(B, cin, H, W, D, generator=g).to(torch.bfloat16).float()
    x1 = torch.randn(B, cin1, H, W, D, generator=g).to(torch.bfloat16).float() if cin1 else None
    xin = x0 if x1 is None else torch.cat((x0, x1), 1)
    if up2:
        xin = F.interpolate(xin, scale_factor=2, mode="nearest")
    ref = F.conv3d(xin, conv.weight.detach(), conv.bias.detach(), stride=stride, padding=k // 2)
    cl = lambda t: None if t is None else t.permute(0, 2, 3, 4, 1).contiguous().cuda().to(torch.bfloat16)
    y, partials, tiles = ops.conv3d(cl(x0), cw.w.cuda(), cw.b.cuda(), cout, k, stride=stride, pad=k // 2, x1=cl(x1),
                                    up2=up2, want_stats=True, out_f32=out_f32, w_tc=cw.w_tc.cuda(),
                                    w_tc_fold=cw.w_tc_fold.cuda() if up2 else None)
    torch.cuda.synchronize()
    got = y.float().permute(0, 4, 1, 2, 3).cpu()
    err = float((got - ref).abs().max() / ref.abs().max())
    V = ref.shape[2] * ref.shape[3] * ref.shape[4]
    stats = ops.instnorm_finalize(partials, V).cpu()
    merr = float((stats[..., 0] - ref.mean(dim=(2, 3, 4))).abs().max())
    tag = f"cin={cin}+{cin1} cout={cout} stride={stride} up2={up2} shape={shape} B={B} {mode}"
    print(f"{'OK ' if err < 1.5e-2 else 'BAD'} {tag}: rel err {err:.3e}, mean err {merr:.2e}, tiles {tiles}", flush=True)
    if err >= 1.5e-2:
        r = ref[0].permute(1, 2, 3, 0).reshape(-1, cout)
        o = got[0].permute(1, 2, 3, 0).reshape(-1, cout)
        print("  ref rows 0..3, cols 0..7:\n", r[:4, :8])
        print("  got rows 0..3, cols 0..7:\n", o[:4, :8])
        bad = ((o - r).abs() > 0.05 * r.abs().max()).float()
        print("  bad fraction per row-block of 8:", bad.mean(1).reshape(-1, 8).mean(1)[:16])
        print("  bad fraction per col-block of 8:", bad.mean(0).reshape(-1, 8).mean(1)[:32])
    return err


def main():
    torch.manual_seed(0)
    print("device:", torch.cuda.get_device_name(0), flush=True)
    run(64, 0, 64, (1, 1, 1), False, (4, 4, 8), "identity")
    run(64, 0, 64, (1, 1, 1), False, (4, 4, 8), "random")
    run(64, 0, 128, (1, 1, 1), False, (4, 4, 8), "random")
    run(16, 0, 16, (1, 1, 1), False, (6, 5, 9), "random")
    run(32, 0, 64, (2, 2, 2), False, (6, 8, 6), "random", B=2)
    run(128, 0, 256, (2, 2, 2), False, (4, 4, 4), "random")
    run(16, 16, 16, (1, 1, 1), False, (6, 6, 5), "random", B=2)
    run(128, 128, 128, (1, 1, 1), False, (3, 4, 4), "random")
    run(128, 0, 32, (1, 1, 1), True, (3, 4, 5), "random")
    run(256, 0, 64, (1, 1, 1), True, (6, 7, 8), "random", B=2)
    run(16, 0, 32, (2, 2, 1), False, (16, 16, 32), "random", B=2)
    run(64, 0, 3, (1, 1, 1), False, (5, 4, 6), "random", B=2, out_f32=True)      # mask head
    run(16, 0, 12, (1, 1, 1), False, (7, 5, 4), "random", B=2, out_f32=True)     # final block
    run(32, 0, 16, (1, 1, 1), False, (5, 6, 7), "random", B=2, k=1)              # gate 1x1x1
    run(8, 0, 16, (1, 1, 1), False, (8, 6, 10), "random", B=2)                   # stem (padded to 8)
    # timing at model shapes (bf16, B=8, 128^3 patch): flops = 2*27*Cin*Cout*Vout
    import time
    def bench(cin, cin1, cout, stride, up2, shape, B=8, k=3, out_f32=False):
        H, W, D = shape
        conv = torch.nn.Conv3d(cin + cin1, cout, k, stride=stride, padding=k // 2)
        cw = _ConvW(conv, want_tc=True, fold_up2=up2)
        x0 = torch.randn(B, H, W, D, cin, device="cuda").to(torch.bfloat16)
        x1 = torch.randn(B, H, W, D, cin1, device="cuda").to(torch.bfloat16) if cin1 else None
        args = (x0, cw.w.cuda(), cw.b.cuda(), cout, k)
        kw = dict(stride=stride, pad=k // 2, x1=x1, up2=up2, want_stats=True, out_f32=out_f32)
        wtc = cw.w_tc.cuda()
        wfold = cw.w_tc_fold.cuda() if up2 else None
        res = {}
        for name, w in (("tc", wtc), ("cuda-core", None)):
            kw["w_tc_fold"] = wfold if name == "tc" else None
            for _ in range(2):
                y, _, _ = ops.conv3d(*args, w_tc=w, **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                y, _, _ = ops.conv3d(*args, w_tc=w, **kw)
            e1.record()
            torch.cuda.synchronize()
            res[name] = e0.elapsed_time(e1) / 5
        V = y.shape[1] * y.shape[2] * y.shape[3]
        fl = 2 * k ** 3 * (cin + cin1) * cout * B * V
        by = (x0.numel() + (0 if x1 is None else x1.numel())) * 2 + y.numel() * y.element_size()
        print(f"time cin={cin}+{cin1} cout={cout} stride={stride} up2={up2} in={shape}: tc {res['tc']:.3f} ms "
              f"({fl/res['tc']/1e9:.1f} TFLOP/s, {by/res['tc']/1e6:.0f} GB/s), cuda-core {res['cuda-core']:.3f} ms ({fl/res['cuda-core']/1e9:.1f} TFLOP/s)",
              flush=True)
    bench(128, 0, 32, (1, 1, 1), True, (39, 23, 64))      # b1.up_embed
    bench(256, 0, 64, (1, 1, 1), True, (24, 14, 32))      # b2.up_embed
    bench(256, 0, 128, (1, 1, 1), True, (15, 9, 32))      # b3.up_embed
    bench(32, 0, 128, (2, 2, 2), False, (78, 46, 128))    # b1.down_embed
    bench(16, 0, 16, (1, 1, 1), False, (64, 64, 128))     # enc.block0.conv1
    bench(16, 0, 32, (2, 2, 1), False, (64, 64, 128))     # enc.block0.conv2
    bench(32, 32, 32, (1, 1, 1), False, (32, 32, 128))    # dec.block2.conv2
    bench(64, 0, 64, (1, 1, 1), False, (16, 16, 64))      # enc.block2.conv1
    bench(128, 0, 256, (2, 2, 2), False, (8, 8, 64))      # enc.block3.conv2
    bench(256, 0, 128, (1, 1, 1), False, (8, 8, 64))      # dec.block0.conv1
    bench(32, 0, 3, (1, 1, 1), False, (64, 64, 128), out_f32=True)     # mask head, finest level
    bench(16, 0, 12, (1, 1, 1), False, (64, 64, 128), out_f32=True)    # final block
    bench(8, 0, 16, (1, 1, 1), False, (64, 64, 128))                   # stem
    bench(32, 0, 16, (1, 1, 1), False, (64, 64, 128), k=1)             # gate W_g
    bench(16, 0, 16, (1, 1, 1), False, (64, 64, 128), k=1)             # gate W_x


if __name__ == "__main__":
    main()
