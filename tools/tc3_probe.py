"""Time the TMA-halo tcgen05 convolution (conv_tc3.cu) against the im2col tcgen05 kernel (conv_tc.cu) on the
model's stride-1 layers with >= 64 input channels (bf16, batch 8 of 128^3 patches), CUDA events."""
import sys
import torch

sys.path.insert(0, ".")
from lintransunet_b200 import ops  # noqa: E402
from lintransunet_b200.unet import _ConvW  # noqa: E402
from tools.ffn_probe import timeit  # noqa: E402

SMALL = [  # the small-channel stride-1 layers (so far: smem-halo mma.sync kernel); (H, W, D) of the conv input
    ("enc.block0.conv1 16->16", 16, 0, 16, False, (64, 64, 128)),
    ("enc.block1.conv1 32->32", 32, 0, 32, False, (32, 32, 128)),
    ("dec.block2.conv2 32+32->32", 32, 32, 32, False, (32, 32, 128)),
    ("dec.block3.conv1 32->16", 32, 0, 16, False, (64, 64, 128)),
    ("dec.block3.conv2 16+16->16", 16, 16, 16, False, (64, 64, 128)),
]
LAYERS = [  # name, cin, cin1, cout, up2, (H, W, D) of the conv INPUT
    ("b1.up_embed 128->32 x2", 128, 0, 32, True, (39, 23, 64)),
    ("b2.up_embed 256->64 x2", 256, 0, 64, True, (24, 14, 32)),
    ("b3.up_embed 256->128 x2", 256, 0, 128, True, (15, 9, 32)),
    ("enc.block2.conv1 64->64", 64, 0, 64, False, (16, 16, 64)),
    ("enc.block3.conv1 128->128", 128, 0, 128, False, (8, 8, 64)),
    ("dec.block0.conv1 256->128", 256, 0, 128, False, (8, 8, 64)),
    ("dec.block0.conv2 128+128->128", 128, 128, 128, False, (8, 8, 64)),
    ("dec.block1.conv1 128->64", 128, 0, 64, False, (16, 16, 64)),
    ("dec.block1.conv2 64+64->64", 64, 64, 64, False, (16, 16, 64)),
    ("dec.block2.conv1 64->32", 64, 0, 32, False, (32, 32, 128)),
]
B = 8
torch.manual_seed(0)
for name, cin, cin1, cout, up2, (H, W, D) in (SMALL + LAYERS if "--all" in sys.argv else (SMALL if "--small" in sys.argv else LAYERS)):
    conv = torch.nn.Conv3d(cin + cin1, cout, 3, padding=1)
    cw = _ConvW(conv, want_tc=True, fold_up2=up2)
    x0 = torch.randn(B, H, W, D, cin, device="cuda").to(torch.bfloat16)
    x1 = torch.randn(B, H, W, D, cin1, device="cuda").to(torch.bfloat16) if cin1 else None
    args = dict(x1=x1, up2=up2, want_stats=True, w_tc=cw.w_tc.cuda(), w_tc_fold=cw.w_tc_fold.cuda() if up2 else None)
    w, b = cw.w.cuda(), cw.b.cuda()
    res = {}
    for flag in (False, True):
        ops.USE_TC3_CONV = flag
        fn = lambda: ops.conv3d(x0, w, b, cout, 3, **args)
        y, part, tiles = fn()
        res[flag] = (timeit(fn), y, ops.instnorm_finalize(part, y.shape[1] * y.shape[2] * y.shape[3]))
    ops.USE_TC3_CONV = True
    dy = (res[True][1].float() - res[False][1].float()).abs().max().item()
    ds = (res[True][2] - res[False][2]).abs().max().item()
    V = B * H * W * D * (8 if up2 else 1)
    fl = 2 * 27 * (cin + cin1) * cout * V
    print(f"{name:32s} before {res[False][0]:8.1f} us | tma-halo {res[True][0]:8.1f} us ({fl / res[True][0] / 1e6:7.0f} TFLOP/s alg.)"
          f" | max|dy| {dy:.4f} max|dstats| {ds:.2e}", flush=True)
