#!/usr/bin/env python
"""Every convolution launch of one bf16 forward (batch 8 of 128^3, 3 classes), in launch order: this repo's kernel against
the same convolution through PyTorch's library path on the same GPU -- F.conv3d in bf16 on the SAME channels-last memory
(`channels_last_3d`, cudnn.benchmark on), including the torch.cat of a two-input convolution and the nn.Upsample(nearest x2)
of the up_embed layers, which the reference executes and this repo folds into the convolution.  Both sides are timed as
CUDA-graph replays with CUDA events.  The library side does not produce the InstanceNorm partial sums nor the fused fp32
mask head; it is a lower bound of what the reference pays for the layer."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lintransunet_b200 import MaskTransUnet, ops  # noqa: E402

CFG = dict(num_layers=[16, 32, 64, 128, 256], roi_size_list=[100, 65, 40, 25, 10],
           is_roi_list=[False, True, True, True, True], dim_input=1, dim_output=3)


def graph_us(fn, reps=5):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    torch.manual_seed(0)
    torch.backends.cudnn.benchmark = True
    m = MaskTransUnet(**CFG).cuda().eval()
    m.use_cuda_graphs = False
    x = torch.randn(8, 1, 128, 128, 128, device="cuda")
    calls = []
    real = ops.conv3d

    def spy(x0, w_packed, bias, cout, ksize, stride=(1, 1, 1), pad=1, x1=None, up2=False, out_f32=False, want_stats=False,
            w_tc=None, w_tc_fold=None, n_aux=0, sv=None):
        calls.append(dict(x0=x0.detach().clone(), x1=None if x1 is None else x1.detach().clone(), w_packed=w_packed, bias=bias,
                          cout=cout, ksize=ksize, stride=tuple(stride), pad=pad, up2=up2, out_f32=out_f32,
                          want_stats=want_stats, w_tc=w_tc, w_tc_fold=w_tc_fold, n_aux=n_aux, sv=sv))
        return real(x0, w_packed, bias, cout, ksize, stride=stride, pad=pad, x1=x1, up2=up2, out_f32=out_f32,
                    want_stats=want_stats, w_tc=w_tc, w_tc_fold=w_tc_fold, n_aux=n_aux, sv=sv)

    with torch.autocast("cuda", dtype=torch.bfloat16):
        m.predict_labels(x)
        ops.conv3d = spy
        m.predict_labels(x)
        ops.conv3d = real
    torch.cuda.synchronize()
    print("| # | layer (channels, spatial, stride) | kernel | ours us | F.conv3d bf16 (cuDNN) us | speed-up |")
    print("|---:|---|---|---:|---:|---:|")
    tot_o = tot_l = 0.0
    for i, c in enumerate(calls):
        x0, x1 = c["x0"], c["x1"]
        B, H, W, D, C0 = x0.shape
        cin = C0 + (0 if x1 is None else x1.shape[-1])
        ctot = c["cout"] + c["n_aux"]
        prof = ops.KernelProfiler()
        ops.set_profiler(prof)
        real(x0, c["w_packed"], c["bias"], c["cout"], c["ksize"], stride=c["stride"], pad=c["pad"], x1=x1, up2=c["up2"],
             out_f32=c["out_f32"], want_stats=c["want_stats"], w_tc=c["w_tc"], w_tc_fold=c["w_tc_fold"], n_aux=c["n_aux"], sv=c["sv"])
        torch.cuda.synchronize()
        ops.set_profiler(None)
        kname = prof.records[0][0] if prof.records else "?"
        t_o = graph_us(lambda: real(x0, c["w_packed"], c["bias"], c["cout"], c["ksize"], stride=c["stride"], pad=c["pad"], x1=x1,
                                    up2=c["up2"], out_f32=c["out_f32"], want_stats=c["want_stats"], w_tc=c["w_tc"],
                                    w_tc_fold=c["w_tc_fold"], n_aux=c["n_aux"], sv=c["sv"]))
        k = c["ksize"]
        stem = C0 == 8 and i == 0                                   # the stem: four real + four zero-padded input channels
        cin_lib = 4 if stem else cin
        w = (torch.randn(ctot, cin_lib, k, k, k, device="cuda") * 0.05).to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
        b = torch.zeros(ctot, device="cuda", dtype=torch.bfloat16)
        xin0 = x0[..., :C0].permute(0, 4, 1, 2, 3)                  # NCDHW view of the same channels-last memory
        if stem:
            xin0 = xin0[:, :4].contiguous(memory_format=torch.channels_last_3d)

        def lib():
            a = xin0 if x1 is None else torch.cat([xin0, x1.permute(0, 4, 1, 2, 3)], 1)
            if c["up2"]:
                a = F.interpolate(a, scale_factor=2, mode="nearest")
            return F.conv3d(a, w, b, stride=c["stride"], padding=c["pad"])
        t_l = graph_us(lib)
        tot_o += t_o; tot_l += t_l
        desc = f"{C0}{'' if x1 is None else '+' + str(x1.shape[-1])}->{c['cout']}{'' if not c['n_aux'] else '+' + str(c['n_aux']) + ' aux'}, " \
               f"{H}x{W}x{D}, k{k}, s{c['stride']}{', nearest x2' if c['up2'] else ''}{', fp32 out' if c['out_f32'] else ''}"
        print(f"| {i} | {desc} | {kname} | {t_o:.1f} | {t_l:.1f} | {t_l / t_o:.2f}x |", flush=True)
    print(f"\nall {len(calls)} convolution launches of one forward: ours {tot_o / 1e3:.2f} ms, library path {tot_l / 1e3:.2f} ms "
          f"({tot_l / tot_o:.2f}x)")


if __name__ == "__main__":
    main()
