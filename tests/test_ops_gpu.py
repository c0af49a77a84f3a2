"""GPU parity of every C-ABI kernel against the CPU oracle (oracle/ltu_oracle.py) and the
reference-generated golden vectors, called through lintransunet_b200.ops (ctypes -> C ABI).

Tolerances (max|a-b| / max|ref|): fp32 storage 1e-4 for the transcendental kernels (ex2-based
exp) and 2e-5 for pure FMA kernels; bf16 storage 1.2e-2 (one bf16 rounding of the output is
2^-9 = 2e-3 relative per element, plus bf16 inputs).  Integer / index results are bit-exact.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ltu_oracle as O
from tests.helpers import load_golden, rel_err, sub

pytestmark = pytest.mark.gpu

DTYPES = [torch.float32, torch.bfloat16]
TOL = {torch.float32: 1e-4, torch.bfloat16: 1.2e-2}


def _ops():
    from lintransunet_b200 import ops
    return ops


def to_cl(t):       # reference [B,C,H,W,D] -> channels-last [B,H,W,D,C]
    return t.permute(0, 2, 3, 4, 1).contiguous()


def from_cl(t):
    return t.permute(0, 4, 1, 2, 3).contiguous()


def rnd(shape, seed, scale=1.0):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale


def q_(t, dtype):   # round through the storage dtype so both sides see the same inputs
    return t.to(dtype).float()


# ------------------------------------------------------------------------------ a1
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("B,h,N", [(2, 4, 100), (1, 8, 333), (1, 8, 31), (3, 4, 4096), (1, 4, 57408 // 4 + 5),
                                   (2, 8, 512), (1, 8, 1), (1, 2, 700), (1, 1, 65)])
def test_attention_core(dtype, B, h, N):
    ops = _ops()
    C = 32 * h
    qkv = q_(rnd((B, N, 3 * C), 10 + N, 2.0), dtype)
    dev = qkv.to("cuda", dtype)
    q, k, v = dev[..., :C], dev[..., C:2 * C], dev[..., 2 * C:]
    ctx = ops.kv_reduce(k, v, h)
    out = ops.q_readout(q, ctx, h)
    split = lambda t: t.view(B, N, h, 32).transpose(1, 2)
    rq, rk, rv = (split(qkv[..., i * C:(i + 1) * C]) for i in range(3))
    ref_ctx = torch.softmax(rk, -2).transpose(-1, -2) @ rv
    ref = O.efficient_attention(rq, rk, rv).transpose(1, 2).reshape(B, N, C)
    # ctx is fp32 in both storage modes; the bf16 path feeds P = exp(k - r) to the tensor pipe in bf16
    assert rel_err(ctx, ref_ctx) < (1e-4 if dtype == torch.float32 else 4e-3)
    assert rel_err(out.float(), ref) < TOL[dtype]
    # determinism: the two-stage reduction has a fixed order
    assert torch.equal(ctx, ops.kv_reduce(k, v, h))


def test_attention_core_golden():
    ops = _ops()
    g = load_golden("ops.npz")
    for tag in ("a", "b"):
        q, k, v = (torch.from_numpy(g[f"attn_{tag}_{n}"]) for n in "qkv")       # [B,h,N,32]
        B, h, N, _ = q.shape
        merge = lambda t: t.transpose(1, 2).reshape(B, N, h * 32).contiguous().cuda()
        out = ops.q_readout(merge(q), ops.kv_reduce(merge(k), merge(v), h), h)
        ref = torch.from_numpy(g[f"attn_{tag}_out"]).transpose(1, 2).reshape(B, N, h * 32)
        assert rel_err(out, ref) < 1e-4


def test_attention_core_large_logits_are_stable():
    """Online max-rescaling: shifted keys/queries must not overflow (softmax is shift invariant)."""
    ops = _ops()
    B, h, N = 1, 4, 1000
    C = 128
    qkv = rnd((B, N, 3 * C), 3, 3.0)
    qkv[..., :2 * C] += 80.0
    qkv[:, 500:, C:2 * C] += 40.0        # the running max changes in the middle of the stream
    dev = qkv.cuda()
    out = ops.q_readout(dev[..., :C], ops.kv_reduce(dev[..., C:2 * C], dev[..., 2 * C:], h), h)
    split = lambda t: t.view(B, N, h, 32).transpose(1, 2).double()
    ref = O.efficient_attention(*(split(qkv[..., i * C:(i + 1) * C]) for i in range(3)))
    ref = ref.transpose(1, 2).reshape(B, N, C)
    assert torch.isfinite(out).all()
    assert rel_err(out, ref) < 2e-4


@pytest.mark.parametrize("h", [4, 8])
def test_attention_core_bf16_reference_rescale_path(h):
    """bf16 tensor-pipe kv_reduce: keys that run 2^64 above the per-CTA reference force the exact
    rescale path; shifted columns must still give the shift-invariant softmax."""
    ops = _ops()
    B, N, C = 2, 3000, 32 * h
    qkv = rnd((B, N, 3 * C), 5, 2.0)
    qkv[:, 1700:, C:2 * C] += 120.0          # exp(120) overflows fp32: only a rescale can survive it
    qkv[:, 2500:, C + 3] += 60.0
    qkv = q_(qkv, torch.bfloat16)
    dev = qkv.to("cuda", torch.bfloat16)
    ctx = ops.kv_reduce(dev[..., C:2 * C], dev[..., 2 * C:], h)
    out = ops.q_readout(dev[..., :C], ctx, h)
    split = lambda t: t.view(B, N, h, 32).transpose(1, 2).double()
    rq, rk, rv = (split(qkv[..., i * C:(i + 1) * C]) for i in range(3))
    ref_ctx = torch.softmax(rk, -2).transpose(-1, -2) @ rv
    ref = O.efficient_attention(rq, rk, rv).transpose(1, 2).reshape(B, N, C)
    assert torch.isfinite(ctx).all() and torch.isfinite(out.float()).all()
    assert rel_err(ctx, ref_ctx) < 4e-3
    assert rel_err(out.float(), ref) < TOL[torch.bfloat16]


# ------------------------------------------------------------------------------ a3 / a4
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("rows,C", [(77, 128), (1000, 256), (1, 256), (4097, 128)])
def test_add_layernorm(dtype, rows, C):
    ops = _ops()
    x, r = q_(rnd((rows, C), 1), dtype), q_(rnd((rows, C), 2), dtype)
    g, b = 1 + 0.1 * rnd((C,), 3), 0.1 * rnd((C,), 4)
    y = ops.add_layernorm(x.to("cuda", dtype), r.to("cuda", dtype), g.cuda(), b.cuda(), 1e-6)
    ref = F.layer_norm(x + r, (C,), g, b, eps=1e-6)
    assert rel_err(y.float(), ref) < (2e-5 if dtype == torch.float32 else TOL[dtype])


@pytest.mark.parametrize("dtype", DTYPES)
def test_gelu(dtype):
    ops = _ops()
    x = q_(rnd((333, 512), 5, 2.0), dtype)
    y = ops.gelu_(x.to("cuda", dtype).clone())
    assert rel_err(y.float(), F.gelu(x)) < (2e-6 if dtype == torch.float32 else TOL[dtype])


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(1, 128, 5, 4, 6), (2, 256, 3, 3, 8), (1, 128, 1, 1, 1), (1, 256, 2, 7, 1),
                                   (1, 128, 5, 4, 37), (2, 256, 3, 4, 16), (1, 128, 2, 2, 64), (2, 128, 9, 7, 70),
                                   (1, 64, 6, 5, 33), (1, 40, 3, 3, 9)])
def test_posenc(dtype, shape):
    ops = _ops()
    from lintransunet_b200.unet import Conv3dPosEmbedding, _pos_w
    B, C, H, W, D = shape
    pe = Conv3dPosEmbedding(C)
    x = q_(rnd(shape, 6), dtype)
    w, b = _pos_w(pe)
    y = ops.posenc_dwconv3(to_cl(x).to("cuda", dtype), w.cuda(), b.cuda())
    ref = O.pos_embedding(x, pe.proj.weight.detach(), pe.proj.bias.detach())
    assert rel_err(from_cl(y.float()), ref) < (2e-5 if dtype == torch.float32 else TOL[dtype])


@pytest.mark.parametrize("rows,C", [(77, 128), (1000, 256), (4097, 256)])
def test_add_layernorm_split_stream(rows, C):
    """hi + lo carries the LayerNorm output to ~16 significant bits; hi alone is the plain bf16 result."""
    ops = _ops()
    bf = torch.bfloat16
    x = rnd((rows, C), 1)
    x_hi = x.to(bf)
    x_lo = (x - x_hi.float()).to(bf)
    r = q_(rnd((rows, C), 2), bf)
    g, b = 1 + 0.1 * rnd((C,), 3), 0.1 * rnd((C,), 4)
    y_hi, y_lo = ops.add_layernorm_split(x_hi.cuda(), x_lo.cuda(), r.to("cuda", bf), g.cuda(), b.cuda(), 1e-6)
    ref = F.layer_norm(x_hi.float() + x_lo.float() + r, (C,), g, b, eps=1e-6)
    assert rel_err(y_hi.float() + y_lo.float(), ref) < 5e-5
    assert torch.equal(y_hi.cpu(), ref.to(bf)) or rel_err(y_hi.float(), ref) < TOL[bf]
    # x_lo = None: the first layer of a stack starts from a plain bf16 tensor
    y2_hi, y2_lo = ops.add_layernorm_split(x_hi.cuda(), None, r.to("cuda", bf), g.cuda(), b.cuda(), 1e-6)
    ref2 = F.layer_norm(x_hi.float() + r, (C,), g, b, eps=1e-6)
    assert rel_err(y2_hi.float() + y2_lo.float(), ref2) < 5e-5
    assert torch.equal(y2_hi, ops.add_layernorm(x_hi.cuda(), r.to("cuda", bf), g.cuda(), b.cuda(), 1e-6))


def test_posenc_split_stream():
    ops = _ops()
    from lintransunet_b200.unet import Conv3dPosEmbedding, _pos_w
    bf = torch.bfloat16
    B, C, H, W, D = 2, 256, 5, 4, 7
    pe = Conv3dPosEmbedding(C)
    x = rnd((B, C, H, W, D), 6)
    x_hi = x.to(bf)
    x_lo = (x - x_hi.float()).to(bf)
    w, b = _pos_w(pe)
    y_hi, y_lo = ops.posenc_dwconv3_split(to_cl(x_hi).cuda(), to_cl(x_lo).cuda(), w.cuda(), b.cuda())
    # the taps read hi (a conv input is bf16 under autocast), the residual term is hi + lo
    conv = O.pos_embedding(x_hi.float(), pe.proj.weight.detach(), pe.proj.bias.detach()) - x_hi.float()
    ref = x_hi.float() + x_lo.float() + conv
    assert rel_err(from_cl(y_hi.float() + y_lo.float()), ref) < 5e-5
    plain = ops.posenc_dwconv3(to_cl(x_hi).cuda(), w.cuda(), b.cuda())
    y0_hi, _ = ops.posenc_dwconv3_split(to_cl(x_hi).cuda(), None, w.cuda(), b.cuda())
    assert torch.equal(y0_hi, plain)


def test_posenc_golden():
    ops = _ops()
    from lintransunet_b200.unet import Conv3dPosEmbedding, _pos_w
    g = load_golden("ops.npz")
    sd = O.make_state_dict(O.UnetConfig(), seed=3)
    pe = Conv3dPosEmbedding(128)
    p = "decode.bridge_list.1.transformer.pos_encoder.proj"
    pe.load_state_dict({"proj.weight": sd[p + ".weight"], "proj.bias": sd[p + ".bias"]})
    w, b = _pos_w(pe)
    y = ops.posenc_dwconv3(to_cl(torch.from_numpy(g["pos_x"])).cuda(), w.cuda(), b.cuda())
    assert rel_err(from_cl(y), g["pos_out"]) < 2e-5


# ------------------------------------------------------------------------------ convolutions
CONV_CASES = [
    # cin, cin1, cout, k, stride, up2, spatial (H,W,D), out_f32
    (4, 0, 16, 3, (1, 1, 1), False, (8, 6, 10), False),        # stem
    (16, 0, 16, 3, (1, 1, 1), False, (6, 5, 9), False),
    (16, 0, 32, 3, (2, 2, 1), False, (8, 6, 7), False),        # DownBlock conv2, stride (2,2,1)
    (32, 0, 64, 3, (2, 2, 2), False, (6, 8, 6), False),
    (128, 0, 256, 3, (2, 2, 2), False, (4, 4, 4), False),
    (64, 0, 2, 3, (1, 1, 1), False, (5, 4, 6), True),          # mask head (fp32 logits)
    (16, 0, 12, 3, (1, 1, 1), False, (7, 5, 4), True),         # final block, 3 classes
    (32, 0, 16, 1, (1, 1, 1), False, (5, 6, 7), False),        # gate W_g 1x1x1
    (16, 16, 16, 3, (1, 1, 1), False, (6, 6, 5), False),       # UpBlock conv2 on cat(x, skip)
    (128, 128, 128, 3, (1, 1, 1), False, (3, 4, 4), False),
    (128, 0, 32, 3, (1, 1, 1), True, (3, 4, 5), False),        # up_embed: nearest x2 folded in
    (32, 0, 128, 3, (2, 2, 2), False, (10, 6, 8), False),      # down_embed
    # TMA halo + tcgen05 kernel (conv_tc3.cu): stride-1 3x3x3, >= 64 channels per input, ragged tiles in every axis
    (64, 0, 64, 3, (1, 1, 1), False, (9, 17, 5), False),
    (256, 0, 128, 3, (1, 1, 1), False, (8, 8, 20), False),      # decoder level 0: D is the 16-axis
    (128, 0, 64, 3, (1, 1, 1), False, (5, 33, 9), False),
    (64, 64, 64, 3, (1, 1, 1), False, (6, 7, 18), False),       # cat(x, skip)
    (64, 0, 32, 3, (1, 1, 1), False, (4, 4, 8), False),         # exactly one tile
    (256, 0, 64, 3, (1, 1, 1), True, (5, 3, 9), False),         # folded up_embed, 2 classes per pass
    (256, 0, 128, 3, (1, 1, 1), True, (3, 5, 4), False),        # folded up_embed, 1 class per pass
    (128, 0, 32, 3, (1, 1, 1), True, (6, 18, 10), False),       # folded up_embed, 4 classes per pass, several tiles
]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("tc", [False, True])
def test_conv3d(dtype, case, tc):
    ops = _ops()
    from lintransunet_b200.unet import _ConvW
    cin, cin1, cout, k, stride, up2, (H, W, D), out_f32 = case
    if tc and dtype != torch.bfloat16:
        pytest.skip("tensor-core path is bf16 only")
    B = 2
    conv = torch.nn.Conv3d(cin + cin1, cout, k, stride=stride, padding=k // 2)
    with torch.no_grad():
        conv.weight.copy_(q_(conv.weight, dtype) if tc else conv.weight)
    cw = _ConvW(conv, want_tc=tc, cin_pad=8 if (tc and cin == 4) else 0, fold_up2=up2)
    x0 = q_(rnd((B, cin, H, W, D), 7), dtype)
    x1 = q_(rnd((B, cin1, H, W, D), 8), dtype) if cin1 else None
    xin = x0 if x1 is None else torch.cat((x0, x1), 1)
    if up2:
        xin = F.interpolate(xin, scale_factor=2, mode="nearest")
    ref = F.conv3d(xin, conv.weight.detach(), conv.bias.detach(), stride=stride, padding=k // 2)
    dev = lambda t: None if t is None else to_cl(t).to("cuda", dtype)
    if tc and cin == 4:                                     # tensor-core stem: input zero-padded to 8 channels
        x0 = torch.cat([x0, torch.zeros_like(x0)], 1)
    y, partials, tiles = ops.conv3d(dev(x0), cw.w.cuda(), cw.b.cuda(), cout, k, stride=stride, pad=k // 2,
                                    x1=dev(x1), up2=up2, out_f32=out_f32, want_stats=True,
                                    w_tc=cw.w_tc.cuda() if tc else None,
                                    w_tc_fold=cw.w_tc_fold.cuda() if (tc and up2) else None)
    assert y.dtype == (torch.float32 if out_f32 else dtype)
    tol = 2e-5 if (dtype == torch.float32 or out_f32) else TOL[torch.bfloat16]   # fp32 output: fp32 math on exact inputs
    assert rel_err(from_cl(y.float()), ref) < tol
    # InstanceNorm statistics come from the fp32 accumulators
    V = ref.shape[2] * ref.shape[3] * ref.shape[4]
    stats = ops.instnorm_finalize(partials, V)
    mean = ref.mean(dim=(2, 3, 4))
    rstd = 1 / torch.sqrt(ref.var(dim=(2, 3, 4), unbiased=False) + 1e-5)
    assert rel_err(stats[..., 0], mean) < 5e-3 and rel_err(stats[..., 1], rstd) < 5e-3
    if not out_f32:
        res = q_(rnd(ref.shape, 9), dtype)
        z = ops.instnorm_apply(y.clone(), stats, ops.ACT_LRELU, residual=to_cl(res).to("cuda", dtype), inplace=True)
        ref_z = F.leaky_relu(F.instance_norm(from_cl(y.float()).cpu(), eps=1e-5), 0.01) + res
        assert rel_err(from_cl(z.float()), ref_z) < (1e-4 if dtype == torch.float32 else TOL[dtype])


SV_CASES = [
    # ci per input, inputs, main channels, fp32 head channels, spatial (H,W,D): super-voxel form on the TMA-halo tcgen05 kernel
    (8, 1, 16, 0, (9, 17, 40)),         # stem (4 real + 4 zero channels): g = 8, UMMA N = 128
    (16, 1, 16, 0, (5, 33, 16)),        # enc.block0.conv1: g = 4, N = 64; ragged tiles
    (16, 1, 16, 0, (64, 64, 32)),       # many tiles, persistent loop wraps
    (32, 1, 32, 0, (9, 5, 18)),         # enc.block1.conv1: g = 2
    (32, 1, 16, 3, (7, 16, 12)),        # dec.block3.conv1 + the finest mask head (fp32 aux)
    (32, 1, 16, 2, (4, 4, 8)),
    (16, 2, 16, 0, (6, 9, 20)),         # dec.block3.conv2 on cat(x, skip)
    (32, 2, 32, 0, (5, 8, 6)),          # dec.block2.conv2 on cat(x, skip)
    (16, 1, 0, 12, (8, 8, 16)),         # final_block, 3 classes: fp32 output only
    (16, 1, 0, 8, (3, 18, 4)),          # final_block, 2 classes
    (32, 1, 0, 3, (6, 6, 10)),          # mask head alone
]


@pytest.mark.parametrize("case", SV_CASES)
def test_conv3d_super_voxel_form(case):
    ops = _ops()
    from lintransunet_b200.unet import _ConvW
    ci, n_in, n_main, n_aux, (H, W, D) = case
    B, bf = 2, torch.bfloat16
    f32_only = n_main == 0
    conv = torch.nn.Conv3d(ci * n_in, n_aux if f32_only else n_main, 3, padding=1)
    aux = None if (f32_only or n_aux == 0) else torch.nn.Conv3d(ci * n_in, n_aux, 3, padding=1)
    with torch.no_grad():
        conv.weight.copy_(q_(conv.weight, bf))
        if aux is not None:
            aux.weight.copy_(q_(aux.weight, bf))
    conv.cuda()
    if aux is not None:
        aux.cuda()
    cw = _ConvW(conv, want_tc=True, aux=aux, n_inputs=n_in, f32_out=f32_only)
    assert cw.sv is not None and cw.sv.g == 64 // ci
    x0 = q_(rnd((B, ci, H, W, D), 70), bf)
    x1 = q_(rnd((B, ci, H, W, D), 71), bf) if n_in == 2 else None
    xin = x0 if x1 is None else torch.cat((x0, x1), 1)
    ref = F.conv3d(xin, conv.weight.detach().cpu(), conv.bias.detach().cpu(), padding=1)
    dev = lambda t: None if t is None else to_cl(t).to("cuda", bf)
    prev_sv, ops.USE_SV_CONV = ops.USE_SV_CONV, True            # opt-in form (slower than conv3d_halo inside the step)
    try:
        res = ops.conv3d(dev(x0), cw.w, cw.b, cw.cout, 3, pad=1, x1=dev(x1), out_f32=f32_only, want_stats=not f32_only,
                         w_tc=cw.w_tc, n_aux=0 if f32_only else n_aux, sv=cw.sv)
    finally:
        ops.USE_SV_CONV = prev_sv
    if f32_only:
        y, partials, tiles = res
        assert y.dtype == torch.float32 and partials is None
        assert rel_err(from_cl(y), ref) < 2e-5
        return
    y, partials = res[0], res[1]
    assert y.dtype == bf and rel_err(from_cl(y.float()), ref) < TOL[bf]
    stats = ops.instnorm_finalize(partials, H * W * D)
    assert rel_err(stats[..., 0], ref.mean(dim=(2, 3, 4))) < 5e-3
    assert rel_err(stats[..., 1], 1 / torch.sqrt(ref.var(dim=(2, 3, 4), unbiased=False) + 1e-5)) < 5e-3
    if n_aux:
        ref_aux = F.conv3d(xin, aux.weight.detach().cpu(), aux.bias.detach().cpu(), padding=1)
        assert res[3].dtype == torch.float32 and rel_err(from_cl(res[3]), ref_aux) < 2e-5
    # the mma.sync halo kernel on the same operands: bf16 storage of the same fp32 sums
    prev_sv, ops.USE_SV_CONV = ops.USE_SV_CONV, False
    try:
        alt = ops.conv3d(dev(x0), cw.w, cw.b, cw.cout, 3, pad=1, x1=dev(x1), want_stats=True, w_tc=cw.w_tc, n_aux=n_aux)
    finally:
        ops.USE_SV_CONV = prev_sv
    assert rel_err(y.float(), alt[0].float()) < 8e-3


HALO_CASES = [
    # cin, cin1, cout, spatial (H,W,D), out_f32        (stride 1, k3: the shared-memory halo / mma.sync kernel)
    (8, 0, 16, (8, 8, 32), False),          # stem (4 real + 4 zero channels)
    (16, 0, 16, (5, 9, 40), False),         # ragged tiles in every axis
    (16, 0, 32, (4, 8, 32), False),
    (32, 0, 16, (4, 4, 32), False),
    (16, 16, 16, (6, 5, 7), False),         # cat(x, skip), tiny depth
    (8, 8, 16, (4, 9, 33), False),
    (32, 0, 32, (7, 3, 35), False),
    (32, 0, 3, (5, 6, 34), True),           # finest mask head, fp32 logits
    (16, 0, 12, (4, 8, 16), True),          # final block, 3 classes
    (16, 0, 2, (3, 3, 3), True),
    (16, 0, 16, (5, 9, 40), False, 1),      # gate W_x, 1x1x1
    (32, 0, 16, (4, 5, 33), False, 1),      # gate W_g, 1x1x1
    (32, 0, 32, (3, 8, 32), False, 1),
    (64, 0, 32, (5, 4, 35), False, 1),      # gate W_g of level 3
    (64, 0, 16, (4, 4, 32), False, 1),
]


@pytest.mark.parametrize("case", HALO_CASES)
def test_conv3d_halo_matches_reference_and_tc_path(case):
    ops = _ops()
    from lintransunet_b200.unet import _ConvW
    cin, cin1, cout, (H, W, D), out_f32 = case[:5]
    k = case[5] if len(case) > 5 else 3
    B = 2
    conv = torch.nn.Conv3d(cin + cin1, cout, k, padding=k // 2)
    with torch.no_grad():
        conv.weight.copy_(q_(conv.weight, torch.bfloat16))
    cw = _ConvW(conv, want_tc=True)
    x0 = q_(rnd((B, cin, H, W, D), 70), torch.bfloat16)
    x1 = q_(rnd((B, cin1, H, W, D), 71), torch.bfloat16) if cin1 else None
    ref = F.conv3d(x0 if x1 is None else torch.cat((x0, x1), 1), conv.weight.detach(), conv.bias.detach(), padding=k // 2)
    dev = lambda t: None if t is None else to_cl(t).to("cuda", torch.bfloat16)
    outs = {}
    # the three bf16 kernels that can run this layer: TMA-halo tcgen05 (default when it applies), smem-halo mma.sync, im2col tcgen05
    import os
    for name, halo, tc3 in (("tc3", True, True), ("halo", True, False), ("im2col", False, False)):
        ops.USE_HALO_CONV, ops.USE_TC3_CONV = halo, tc3
        os.environ["LTU_TC3_SMALL"] = "1" if tc3 else "0"       # let the TMA-halo kernel take 16- / 32-channel K slices
        try:
            y, partials, tiles = ops.conv3d(dev(x0), cw.w.cuda(), cw.b.cuda(), cout, k, pad=k // 2, x1=dev(x1),
                                            out_f32=out_f32, want_stats=True, w_tc=cw.w_tc.cuda())
        finally:
            ops.USE_HALO_CONV, ops.USE_TC3_CONV = True, True
            os.environ["LTU_TC3_SMALL"] = "0"
        assert rel_err(from_cl(y.float()), ref) < (2e-5 if out_f32 else TOL[torch.bfloat16]), name
        stats = ops.instnorm_finalize(partials, H * W * D)
        assert rel_err(stats[..., 0], ref.mean(dim=(2, 3, 4))) < 5e-3, name
        assert rel_err(stats[..., 1], 1 / torch.sqrt(ref.var(dim=(2, 3, 4), unbiased=False) + 1e-5)) < 5e-3, name
        outs[name] = y.float()
    assert rel_err(outs["halo"], outs["im2col"]) < (2e-5 if out_f32 else 8e-3)
    assert rel_err(outs["tc3"], outs["im2col"]) < (2e-5 if out_f32 else 8e-3)


@pytest.mark.parametrize("cin,cmain,naux,shape", [(32, 16, 3, (5, 6, 34)), (64, 32, 3, (4, 5, 9)), (256, 128, 2, (3, 4, 4)),
                                                   (128, 64, 3, (4, 4, 6)), (16, 16, 3, (4, 8, 32))])
def test_conv3d_fused_mask_head(cin, cmain, naux, shape):
    """UpBlock.conv1 and the mask head read the same input: one launch, bf16 main output (+ IN statistics)
    and unrounded fp32 logits as the auxiliary output (halo kernel for Cin<=32, tcgen05 otherwise)."""
    ops = _ops()
    from lintransunet_b200.unet import _ConvW
    H, W, D = shape
    B = 2
    main, head = torch.nn.Conv3d(cin, cmain, 3, padding=1), torch.nn.Conv3d(cin, naux, 3, padding=1)
    with torch.no_grad():
        main.weight.copy_(q_(main.weight, torch.bfloat16))
        head.weight.copy_(q_(head.weight, torch.bfloat16))
    cw = _ConvW(main, True, aux=head)
    assert cw.cout == cmain and cw.n_aux == naux
    x = q_(rnd((B, cin, H, W, D), 60), torch.bfloat16)
    ref_main = F.conv3d(x, main.weight.detach(), main.bias.detach(), padding=1)
    ref_head = F.conv3d(x, head.weight.detach(), head.bias.detach(), padding=1)
    y, partials, tiles, aux = ops.conv3d(to_cl(x).to("cuda", torch.bfloat16), cw.w.cuda(), cw.b.cuda(), cmain, 3, pad=1,
                                         want_stats=True, w_tc=cw.w_tc.cuda(), n_aux=naux)
    assert y.shape[-1] == cmain and aux.shape[-1] == naux and aux.dtype == torch.float32
    assert rel_err(from_cl(y.float()), ref_main) < TOL[torch.bfloat16]
    assert rel_err(from_cl(aux), ref_head) < 2e-5
    stats = ops.instnorm_finalize(partials, H * W * D)
    assert rel_err(stats[..., 0], ref_main.mean(dim=(2, 3, 4))) < 5e-3


@pytest.mark.parametrize("dtype", DTYPES)
def test_chan_stats(dtype):
    ops = _ops()
    x = q_(rnd((2, 32, 9, 7, 11), 12) * 3 + 1.5, dtype)
    stats = ops.chan_stats(to_cl(x).to("cuda", dtype))
    assert rel_err(stats[..., 0], x.mean(dim=(2, 3, 4))) < 1e-5
    assert rel_err(stats[..., 1], 1 / torch.sqrt(x.var(dim=(2, 3, 4), unbiased=False) + 1e-5)) < 1e-5


# ------------------------------------------------------------------------------ plumbing
@pytest.mark.parametrize("dtype", DTYPES)
def test_s2d_input(dtype):
    ops = _ops()
    x = rnd((2, 1, 8, 12, 5), 13)
    y = ops.s2d_input(x.cuda(), dtype)
    assert torch.equal(from_cl(y.float()).cpu(), O.space_to_depth(x).to(dtype).float())
    y8 = ops.s2d_input(x.cuda(), dtype, cpad=8)
    assert torch.equal(y8[..., :4], y) and float(y8[..., 4:].abs().max()) == 0.0


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("fd", [1, 2])
@pytest.mark.parametrize("shape", [(2, 16, 3, 4, 5), (1, 256, 1, 1, 4), (1, 32, 4, 4, 16)])
def test_upsample_trilinear(dtype, fd, shape):
    ops = _ops()
    x = q_(rnd(shape, 14), dtype)
    y = ops.upsample_trilinear(to_cl(x).to("cuda", dtype), fd)
    ref = O.trilinear_up(x, (2, 2, fd))
    assert rel_err(from_cl(y.float()), ref) < (2e-6 if dtype == torch.float32 else TOL[dtype])


@pytest.mark.parametrize("cout", [2, 3])
def test_mask_softmax(cout):
    ops = _ops()
    l = rnd((2, cout, 5, 6, 7), 15, 2.0)
    mask, fg = ops.mask_softmax(to_cl(l).cuda(), want_mask=True)
    ref = torch.softmax(l, 1)
    assert rel_err(mask, ref) < 2e-6
    assert rel_err(fg, 1 - ref[:, 0]) < 2e-6


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("ci,cu", [(16, 32), (128, 256), (64, 128)])
def test_attention_gate(dtype, ci, cu):
    ops = _ops()
    from lintransunet_b200.unet import SpatialAttention3DBlock, _ConvW
    blk = SpatialAttention3DBlock(ci, cu, ci)
    skip, up = q_(rnd((2, ci, 5, 4, 6), 16), dtype), q_(rnd((2, cu, 5, 4, 6), 17), dtype)
    sd = {"g." + k: v.detach() for k, v in blk.state_dict().items()}
    ref = skip * O.attention_gate(skip, up, sd, "g")
    wx, wg = _ConvW(blk.W_x[0], False), _ConvW(blk.W_g[0], False)
    s_cl, u_cl = to_cl(skip).to("cuda", dtype), to_cl(up).to("cuda", dtype)
    a, pa, _ = ops.conv3d(s_cl, wx.w.cuda(), wx.b.cuda(), ci, 1, pad=0, want_stats=True)
    g, pg, _ = ops.conv3d(u_cl, wg.w.cuda(), wg.b.cuda(), ci, 1, pad=0, want_stats=True)
    V = 5 * 4 * 6
    out = ops.gate_fused(a, ops.instnorm_finalize(pa, V), g, ops.instnorm_finalize(pg, V),
                         blk.psi[0].weight.detach().reshape(-1).cuda(), blk.psi[0].bias.detach().cuda(), s_cl)
    assert rel_err(from_cl(out.float()), ref) < (1e-4 if dtype == torch.float32 else 2.5e-2)


def _golden_masks(g):
    shape = tuple(int(s) for s in g["box_masks_shape"])
    bits = np.unpackbits(g["box_masks_packed"])[: int(np.prod(shape))]
    return torch.from_numpy(bits.reshape(shape).astype(np.float32))


def test_roi_bbox_bit_exact():
    ops = _ops()
    g = load_golden("ops.npz")
    masks = _golden_masks(g)                                   # [5,1,96,96,8]
    rc = O.UnetConfig().roi_consts(1)
    box = ops.roi_bbox(masks[:, 0].contiguous().cuda(), rc["min_h"], rc["min_w"], 0.5)
    assert np.array_equal(box.cpu().numpy(), g["box_out"])
    # random soft masks at several plane sizes, including degenerate (S < min_roi) ones
    for i, (h, w, d) in enumerate([(16, 16, 8), (32, 24, 16), (128, 96, 4), (7, 5, 3)]):
        fg = torch.rand(3, 1, h, w, d, generator=torch.Generator().manual_seed(20 + i)) ** (i + 1)
        fg[1] = 0
        for lvl in (1, 2, 3):
            rc = O.UnetConfig().roi_consts(lvl)
            ref = O.roi_boxes(fg, rc["min_h"], rc["min_w"])
            got = ops.roi_bbox(fg[:, 0].contiguous().cuda(), rc["min_h"], rc["min_w"], 0.5)
            assert np.array_equal(got.cpu().numpy(), ref.numpy()), (h, w, d, lvl)


@pytest.mark.parametrize("dtype", DTYPES)
def test_roi_resample_golden_and_oracle(dtype):
    ops = _ops()
    g = load_golden("ops.npz")
    rc = O.UnetConfig().roi_consts(1)
    box = torch.from_numpy(g["box_out"])
    feat = torch.randn(5, 4, 96, 96, 8, generator=torch.Generator().manual_seed(int(g["resample_feat_seed"])))
    feat8 = q_(torch.cat([feat, feat * 0.5], 1), dtype)        # 8 channels: one bf16 vector
    geo = (rc["h_roi"], rc["w_roi"], rc["eval_h"], rc["eval_w"])
    roi = ops.roi_resample(to_cl(feat8).to("cuda", dtype), box.cuda(), (96, 96), *geo, direction=0)
    back = ops.roi_resample(roi, box.cuda(), (96, 96), *geo, direction=1)
    x0, y0, x1, y1 = box[:, 0:1], box[:, 1:2], box[:, 3:4], box[:, 4:5]
    ref_roi = O.separable_resample(feat8, O.fisheye_forward_coords(x0, x1, 95, geo[0], geo[2]),
                                   O.fisheye_forward_coords(y0, y1, 95, geo[1], geo[3]))
    ref_back = O.separable_resample(from_cl(roi.float()).cpu(), O.fisheye_back_coords(x0, x1, 95, geo[0], geo[2]),
                                    O.fisheye_back_coords(y0, y1, 95, geo[1], geo[3]))
    tol = 5e-5 if dtype == torch.float32 else TOL[dtype]
    assert rel_err(from_cl(roi.float()), ref_roi) < tol
    assert rel_err(from_cl(back.float()), ref_back) < tol
    if dtype == torch.float32:                                 # against the unmodified reference
        assert rel_err(sub(from_cl(roi)[:, :4], 65536), g["resample_roi"]) < 5e-5
        assert rel_err(sub(from_cl(back)[:, :4], 65536), g["resample_back"]) < 5e-5


def test_roi_resample_degenerate_box_is_all_zero():
    """at the BASELINE patch shapes x0 > x1 and every sample falls outside the map."""
    ops = _ops()
    rc = O.UnetConfig().roi_consts(3)
    box = torch.tensor([[6.5, 3.5, 0.0, -4.5, -1.5, 7.0]])
    x = rnd((1, 4, 4, 8, 128), 31).cuda()
    roi = ops.roi_resample(x, box.cuda(), (4, 4), rc["h_roi"], rc["w_roi"], rc["eval_h"], rc["eval_w"], 0)
    assert roi.shape == (1, 30, 18, 8, 128) and float(roi.abs().max()) == 0.0


@pytest.mark.parametrize("cout", [2, 3])
def test_head(cout):
    ops = _ops()
    logits = rnd((2, 4 * cout, 6, 5, 7), 40, 1.5)
    probs, _, _ = ops.head_d2s_softmax(to_cl(logits).cuda(), cout, True, False, False)
    _, onehot, labels = ops.head_d2s_softmax(to_cl(logits).cuda(), cout, False, True, True)
    ref = torch.softmax(O.depth_to_space(logits), 1)
    assert rel_err(probs, ref) < 2e-6
    idx = ref.argmax(1)
    assert float((labels.cpu().long() != idx).float().mean()) < 1e-4
    assert torch.equal(onehot.argmax(1).cpu(), labels.cpu().long()) and float(onehot.sum(1).min()) == 1.0


def test_vote_accumulate_and_argmax():
    ops = _ops()
    g = torch.Generator().manual_seed(50)
    C, H, W, D, r = 3, 12, 16, 8, 8
    starts = torch.tensor([[0, 0, 0], [4, 0, 0], [4, 8, 0], [0, 8, 0], [2, 4, 0]], dtype=torch.int32)
    labels = torch.randint(0, C, (5, r, r, r), generator=g, dtype=torch.uint8)
    votes = torch.zeros(C, H, W, D, dtype=torch.uint8, device="cuda")
    ops.vote_accumulate(labels.cuda(), starts.cuda(), votes)
    ref = np.zeros((C, H, W, D), np.int64)
    for n in range(5):
        h0, w0, d0 = starts[n].tolist()
        for c in range(C):
            ref[c, h0:h0 + r, w0:w0 + r, d0:d0 + r] += (labels[n].numpy() == c)
    assert np.array_equal(votes.cpu().numpy().astype(np.int64), ref)
    assert np.array_equal(ops.vote_argmax(votes).cpu().numpy(), ref.argmax(0).astype(np.uint8))


# ------------------------------------------------------------------------------ block level
@pytest.mark.parametrize("dtype", DTYPES)
def test_encoder_layer_golden(dtype):
    from lintransunet_b200 import MaskTransUnet
    from lintransunet_b200.unet import SelfAttentionLayer, _LayerW
    g = load_golden("ops.npz")
    sd = O.make_state_dict(O.UnetConfig(), seed=3)
    pre = "decode.bridge_list.1.transformer.layers.2"
    layer = SelfAttentionLayer(128, 4)
    layer.load_state_dict({k[len(pre) + 1:]: v for k, v in sd.items() if k.startswith(pre + ".")})
    layer.cuda()
    x = torch.from_numpy(g["layer_x"])
    for fused_linear, fused_ffn, fused_attn in ((True, False, False), (False, True, False), (False, True, True),
                                                (False, False, True), (False, False, False)):
        m = MaskTransUnet.__new__(MaskTransUnet)      # only _encoder_layer is exercised
        torch.nn.Module.__init__(m)
        m.use_fused_linear, m.use_fused_ffn, m.use_fused_attn = fused_linear, fused_ffn, fused_attn
        # K/V half of the fused_attn paths: ltu_kv_project_reduce | ltu_linear_fused + kv_reduce | cuBLAS + kv_reduce
        for native_linear, fuse_kv, fold in ((True, True, True), (True, True, False), (True, False, True), (True, False, False),
                                             (False, False, False)):
            m.use_native_linear, m.fuse_kv_project, m.fold_readout = native_linear, fuse_kv, fold and dtype == torch.bfloat16
            y, lo = MaskTransUnet._encoder_layer(m, x.to("cuda", dtype), _LayerW(layer, dtype))
            assert lo is None
            assert rel_err(y.float(), g["layer_out"]) < (1e-4 if dtype == torch.float32 else 3e-2), (fused_linear, fused_ffn, fused_attn)
    if dtype == torch.bfloat16:     # split token stream (separate-kernel path): hi + lo carries 16 significant bits
        m.use_fused_linear = m.use_fused_ffn = m.use_fused_attn = False
        xin = x.to("cuda", dtype)
        hi, lo = MaskTransUnet._encoder_layer(m, xin, _LayerW(layer, dtype), None, True)
        plain, _ = MaskTransUnet._encoder_layer(m, xin, _LayerW(layer, dtype))
        assert lo is not None
        e_split, e_plain = rel_err(hi.float() + lo.float(), g["layer_out"]), rel_err(plain.float(), g["layer_out"])
        print(f"\n[encoder layer bf16] plain stream {e_plain:.2e}, split stream (hi + lo) {e_split:.2e}")
        assert e_split <= e_plain and e_split < 3e-2


@pytest.mark.parametrize("B,N", [(1, 300), (2, 4320), (8, 512)])
def test_encoder_layer_d256_native_linear(B, N):
    """d_model-256 SelfAttentionLayer (bridges 2-4, trans_block.py:203-211) with every Linear on ltu_linear_fused
    (bias / GELU / residual + LayerNorm in the GEMM epilogues, split token stream through the operand ring) against
    the oracle in fp64 on the same bf16-rounded input, and against the cuBLAS + separate-kernel path."""
    from lintransunet_b200 import MaskTransUnet
    from lintransunet_b200.unet import SelfAttentionLayer, _LayerW
    bf = torch.bfloat16
    sd = O.make_state_dict(O.UnetConfig(), seed=5)
    pre = "decode.bridge_list.2.transformer.layers.1"
    sub = {k[len(pre) + 1:]: v for k, v in sd.items() if k.startswith(pre + ".")}
    layer = SelfAttentionLayer(256, 8)
    layer.load_state_dict(sub)
    layer.cuda()
    x = q_(rnd((B, N, 256), 95, 1.0), bf)
    x_lo = q_(rnd((B, N, 256), 96, 2.0 ** -10), bf)
    # oracle on the exact split input: the Linear layers see x_hi (autocast casts a Linear's input), the residual adds x_hi + x_lo
    sdd = {pre + "." + k: v.double() for k, v in sub.items()}
    ref_plain = O.encoder_layer(x.double(), sdd, pre, 8)
    m = MaskTransUnet.__new__(MaskTransUnet)
    torch.nn.Module.__init__(m)
    m.use_fused_linear, m.use_fused_ffn, m.use_fused_attn, m.fuse_kv_project = False, False, False, True
    m.fold_readout = True
    lw = _LayerW(layer, bf)
    assert lw.lin
    out = {}
    for native in (True, "separate readout", False):
        m.use_native_linear = bool(native)
        m.fold_readout = native is True
        hi, lo = MaskTransUnet._encoder_layer(m, x.to("cuda", bf), lw, None, True)
        assert lo is not None
        out[native] = (hi, lo)
        assert rel_err(hi.float() + lo.float(), ref_plain) < 2.5e-2, native
        hi1, lo1 = MaskTransUnet._encoder_layer(m, x.to("cuda", bf), lw, None, False)
        assert lo1 is None and rel_err(hi1.float(), ref_plain) < 3e-2
    e_nat = rel_err(out[True][0].float() + out[True][1].float(), ref_plain)
    e_lib = rel_err(out[False][0].float() + out[False][1].float(), ref_plain)
    e_sep = rel_err(out["separate readout"][0].float() + out["separate readout"][1].float(), ref_plain)
    print(f"\n[d256 layer B={B} N={N}] native Linear, readout folded into the GEMMs {e_nat:.2e}, native Linear + q_readout "
          f"{e_sep:.2e}, cuBLAS + separate kernels {e_lib:.2e} (vs fp64 oracle)")
    assert e_nat <= 1.5 * e_lib + 1e-3 and e_nat <= 1.5 * e_sep + 1e-3
    m.fold_readout = True
    # with a non-zero low word on the input (layers 1..7 of a bridge)
    m.use_native_linear = True
    hi2, lo2 = MaskTransUnet._encoder_layer(m, x.to("cuda", bf), lw, x_lo.to("cuda", bf), True)
    m.use_native_linear = False
    hi3, lo3 = MaskTransUnet._encoder_layer(m, x.to("cuda", bf), lw, x_lo.to("cuda", bf), True)
    assert rel_err(hi2.float() + lo2.float(), hi3.float() + lo3.float()) < 2e-2


@pytest.mark.parametrize("epi", [0, 1, 2])
@pytest.mark.parametrize("rows,cin,cout", [(1000, 128, 128), (777, 256, 512), (300, 512, 256), (129, 128, 256),
                                           (40000, 256, 256), (1, 256, 768), (5000, 64, 32)])
def test_linear_tc_fused_epilogues(epi, rows, cin, cout):
    """Persistent tcgen05 Linear layer: bias | bias+GELU(erf) | residual+LayerNorm epilogues vs fp32 torch."""
    ops = _ops()
    if epi == 2 and (cout > 256 or cout % 32):
        pytest.skip("the LayerNorm epilogue needs the whole row in one tile")
    lin = torch.nn.Linear(cin, cout)
    with torch.no_grad():
        lin.weight.copy_(q_(lin.weight, torch.bfloat16))
    x = q_(rnd((rows, cin), 80), torch.bfloat16)
    res = q_(rnd((rows, cout), 81), torch.bfloat16)
    g, b = 1 + 0.1 * rnd((cout,), 82), 0.1 * rnd((cout,), 83)
    ref = F.linear(x, lin.weight.detach(), lin.bias.detach())
    if epi == 1:
        ref = F.gelu(ref)
    elif epi == 2:
        ref = F.layer_norm(ref + res, (cout,), g, b, eps=1e-6)
    y = ops.linear_tc(x.to("cuda", torch.bfloat16), ops.pack_linear_tc(lin.weight).cuda(), lin.bias.detach().float().cuda(),
                      cout, epi, residual=res.to("cuda", torch.bfloat16) if epi == 2 else None,
                      gamma=g.cuda() if epi == 2 else None, beta=b.cuda() if epi == 2 else None)
    assert y.shape == (rows, cout)
    assert rel_err(y.float(), ref) < TOL[torch.bfloat16]


@pytest.mark.parametrize("epi", [0, 1, 2, 3])
@pytest.mark.parametrize("rows,K,N", [(1000, 256, 256), (777, 256, 512), (300, 512, 256), (129, 128, 256), (1, 256, 768),
                                      (40000, 256, 768), (148 * 128 * 2 + 77, 512, 256), (5000, 64, 256), (4096, 256, 512)])
def test_linear_fused(epi, rows, K, N):
    """TMA + tcgen05 Linear (csrc/linear_tma.cu): bias | bias + GELU(erf) | residual + LayerNorm on the split stream
    (epi 3 = the same with res_lo absent and y_lo not requested) vs fp64 torch on the same bf16 operands."""
    ops = _ops()
    ln = epi >= 2
    if ln and N != 256:
        pytest.skip("the LayerNorm epilogue needs the whole 256-wide row in one tile")
    lin = torch.nn.Linear(K, N)
    with torch.no_grad():
        lin.weight.copy_(q_(lin.weight * 2, torch.bfloat16))
    x = q_(rnd((rows, K), 80 + rows % 5, 1.5), torch.bfloat16)
    r_hi = q_(rnd((rows, N), 81, 2.0) + 0.7, torch.bfloat16)
    r_lo = q_(rnd((rows, N), 84, 2.0 ** -9), torch.bfloat16)
    g, b = 1 + 0.1 * rnd((N,), 82), 0.1 * rnd((N,), 83)
    ref = F.linear(x.double(), lin.weight.detach().double(), lin.bias.detach().double())
    if epi == 1:
        ref = F.gelu(ref)
    elif ln:
        ref = F.layer_norm(ref + r_hi.double() + (r_lo.double() if epi == 2 else 0), (N,), g.double(), b.double(), eps=1e-6)
    cu = lambda t: t.detach().to("cuda")
    bf = torch.bfloat16
    out = ops.linear_fused(cu(x).to(bf), cu(lin.weight).to(bf), cu(lin.bias).float(), min(epi, 2),
                           res_hi=cu(r_hi).to(bf) if ln else None, res_lo=cu(r_lo).to(bf) if epi == 2 else None,
                           gamma=cu(g) if ln else None, beta=cu(b) if ln else None, want_lo=(epi == 2))
    if ln:
        y, y_lo = out
        assert (y_lo is None) == (epi == 3)
        assert y.shape == (rows, N)
        assert rel_err(y.float().cpu().double(), ref) < TOL[bf]
        if y_lo is not None:        # hi + lo carries ~16 significant bits of the fp32 LayerNorm output
            assert rel_err(y.double().cpu() + y_lo.double().cpu(), ref) < 2e-4
            assert float((y_lo.float().abs() / y.float().abs().clamp_min(1e-30)).max()) <= 2.0 ** -8   # lo = rounding error of hi
    else:
        assert out.shape == (rows, N)
        assert rel_err(out.float().cpu().double(), ref) < TOL[bf]


@pytest.mark.parametrize("B,N", [(1, 300), (3, 4320), (8, 512), (2, 10752), (5, 97)])
def test_linear_fused_folded_readout(B, N):
    """The query half of linear_attention (trans_block.py:50, :65) carried by the GEMMs around it: (a) the QKV launch writes
    softmax(Q) / sqrt(32) per head in its first C columns, K and V unchanged; (b) ctx_project builds the per-sample weight
    W_b = blockdiag(ctx_b) Wo^T; (c) the output projection with per-sample weights, reading P in place inside the QKV rows,
    equals q_readout followed by the projection -- against fp64 torch on the same bf16 operands, with sample-aligned row
    tiles (N not a multiple of 128: nothing may leak across samples)."""
    ops = _ops()
    bf, C, h = torch.bfloat16, 256, 8
    cu = lambda t: t.detach().to("cuda")
    x = q_(rnd((B, N, C), 180 + N % 7, 1.5), bf)
    w = q_(rnd((3 * C, C), 181, 0.08), bf)
    b = rnd((3 * C,), 182, 0.3)
    qkv_ref = x.double() @ w.double().t() + b.double()
    p_ref = torch.softmax(qkv_ref[..., :C].reshape(B, N, h, 32), dim=-1).reshape(B, N, C) / math.sqrt(32.0)
    qkv = ops.linear_fused(cu(x).to(bf), cu(w).to(bf), cu(b), softmax_cols=C)
    assert qkv.shape == (B, N, 3 * C)
    plain = ops.linear_fused(cu(x).to(bf), cu(w).to(bf), cu(b))
    assert torch.equal(qkv[..., C:], plain[..., C:])                            # K and V are untouched
    assert rel_err(qkv[..., :C].float().cpu().double(), p_ref) < TOL[bf]
    # (b) per-sample weight
    ctx = rnd((B, h, 32, 32), 183, 1.0)
    wo = q_(rnd((C, C), 184, 0.1), bf)
    bo = rnd((C,), 185, 0.2)
    wb = ops.ctx_project(cu(ctx), cu(wo).to(bf))
    wb_ref = torch.einsum("bhje,nhe->bnhj", ctx.double(), wo.double().reshape(C, h, 32)).reshape(B, C, C)
    assert wb.shape == (B, C, C) and rel_err(wb.float().cpu().double(), wb_ref) < TOL[bf]
    # (c) output projection of P with W_b + residual + LayerNorm == readout, projection, residual, LayerNorm
    r_hi = q_(rnd((B, N, C), 186, 2.0) + 0.5, bf)
    r_lo = q_(rnd((B, N, C), 187, 2.0 ** -9), bf)
    g, be = 1 + 0.1 * rnd((C,), 188), 0.1 * rnd((C,), 189)
    P = qkv[..., :C].float().cpu().double()
    att = torch.einsum("bnhj,bhje->bnhe", P.reshape(B, N, h, 32), ctx.double()).reshape(B, N, C)
    ref = F.layer_norm(att @ wo.double().t() + bo.double() + r_hi.double() + r_lo.double(), (C,), g.double(), be.double(), eps=1e-6)
    y, y_lo = ops.linear_fused(qkv, wb, cu(bo), ops.EPI_RES_LN, cu(r_hi).to(bf), cu(r_lo).to(bf), cu(g), cu(be), 1e-6,
                               want_lo=True, x_cols=C)
    assert y.shape == (B, N, C) and torch.isfinite(y.float()).all()
    assert rel_err(y.float().cpu().double(), ref) < TOL[bf]
    assert rel_err(y.double().cpu() + y_lo.double().cpu(), ref) < 3e-3          # W_b is rounded to bf16 once per sample
    # every sample alone gives the same bits (a sample's rows never see a neighbour's weight or rows)
    for s_ in (0, B - 1):
        y1, _ = ops.linear_fused(qkv[s_:s_ + 1].contiguous(), wb[s_:s_ + 1].contiguous(), cu(bo), ops.EPI_RES_LN,
                                 cu(r_hi[s_:s_ + 1]).to(bf), cu(r_lo[s_:s_ + 1]).to(bf), cu(g), cu(be), 1e-6, want_lo=True, x_cols=C)
        assert torch.equal(y1[0], y[s_])
    with pytest.raises(RuntimeError):
        ops.linear_fused(cu(x).to(bf), cu(w).to(bf), cu(b), softmax_cols=100)
    if B > 1:
        with pytest.raises(ValueError):
            ops.linear_fused(qkv[:1].contiguous(), wb, cu(bo), ops.EPI_RES_LN, cu(r_hi).to(bf), None, cu(g), cu(be), x_cols=C)
    # the merge kernel of kv_reduce writes the same W_b for the context it produces
    ctx2, wb2 = ops.kv_reduce(qkv[..., C:2 * C], qkv[..., 2 * C:], h, w_o=cu(wo).to(bf))
    assert torch.equal(ctx2, ops.kv_reduce(qkv[..., C:2 * C], qkv[..., 2 * C:], h))
    assert torch.equal(wb2, ops.ctx_project(ctx2, cu(wo).to(bf)))


def test_concat2_and_the_concatenated_tc3_path():
    """ltu_concat2 == torch.cat on channels-last rows; a 3x3x3 layer with two 32-channel inputs (dec.block2.conv2) runs as ONE
    64-channel input of the TMA-halo tcgen05 kernel and gives the bits of ... the same exact bf16 products, fp32 accumulation
    (summation order differs from the im2col kernel: compared to fp64)."""
    ops = _ops()
    bf = torch.bfloat16
    a = rnd((2, 6, 10, 16, 32), 300).to("cuda", bf)
    b = rnd((2, 6, 10, 16, 32), 301).to("cuda", bf)
    assert torch.equal(ops.concat2(a, b), torch.cat([a, b], -1))
    af = rnd((3, 5, 8), 302).cuda()
    assert torch.equal(ops.concat2(af, rnd((3, 5, 4), 303).cuda())[..., :8], af)
    from lintransunet_b200.unet import _ConvW
    conv = torch.nn.Conv3d(64, 32, 3, padding=1)
    with torch.no_grad():
        conv.weight.copy_(q_(conv.weight, bf))
    conv.cuda()
    cw = _ConvW(conv, want_tc=True, n_inputs=2)
    xin = torch.cat([a, b], -1).float().permute(0, 4, 1, 2, 3).double().cpu()
    ref = F.conv3d(xin, conv.weight.detach().double().cpu(), conv.bias.detach().double().cpu(), padding=1).permute(0, 2, 3, 4, 1)
    outs = {}
    for knob in (True, False):
        ops.USE_CONCAT_TC3 = knob
        try:
            y, part, _ = ops.conv3d(a, cw.w, cw.b, cw.cout, 3, pad=1, x1=b, want_stats=True, w_tc=cw.w_tc)
        finally:
            ops.USE_CONCAT_TC3 = True
        outs[knob] = y
        assert rel_err(y.float().cpu().double(), ref) < TOL[bf]
        V = y.shape[1] * y.shape[2] * y.shape[3]
        st = ops.instnorm_finalize(part, V)
        assert rel_err(st[..., 0].cpu().double(), y.float().cpu().double().reshape(2, V, 32).mean(1)) < 1e-3
    assert rel_err(outs[True].float(), outs[False].float()) < 1e-2


def test_linear_fused_is_deterministic_and_rejects_bad_shapes():
    ops = _ops()
    bf = torch.bfloat16
    x = rnd((5000, 256), 85).to("cuda", bf)
    w = rnd((768, 256), 86, 0.1).to("cuda", bf)
    b = rnd((768,), 87).cuda()
    assert torch.equal(ops.linear_fused(x, w, b), ops.linear_fused(x, w, b))
    with pytest.raises(RuntimeError):
        ops.linear_fused(x, w[:200].contiguous(), b[:200].contiguous())        # N not a multiple of 256
    with pytest.raises(RuntimeError):
        ops.linear_fused(x, w, b, ops.EPI_RES_LN, res_hi=x)                     # LayerNorm epilogue needs N == 256


@pytest.mark.parametrize("B,N", [(1, 32), (1, 96), (2, 160), (3, 4096), (8, 57408), (5, 7008), (2, 64 * 897), (8, 2976)])
def test_kv_project_reduce(B, N):
    """K/V projection + context reduction in one launch (csrc/kv_project2.cu: G = P^T x on the tensor pipe, K and V never
    rounded to bf16) against (a) fp64 torch on the same bf16 x and weights and (b) the separate native path linear_fused ->
    kv_reduce, which rounds K and V to bf16."""
    ops = _ops()
    bf = torch.bfloat16
    C, h = 128, 4
    g = torch.Generator().manual_seed(100 + N % 13)
    x = (torch.randn(B, N, C, generator=g) * 1.3).to("cuda", bf)
    w = (torch.randn(2 * C, C, generator=g) * 0.12).to("cuda", bf)
    b = (torch.randn(2 * C, generator=g) * 0.3).cuda()
    assert ops.kv_project_reduce_supported(C, h, N)
    ctx = ops.kv_project_reduce(x, w, b, h)
    kv = ops.linear_fused(x, w, b)
    ref_native = ops.kv_reduce(kv[..., :C], kv[..., C:], h)
    assert ctx.shape == (B, h, 32, 32)
    kv64 = x.double() @ w.double().t() + b.double()
    k64 = kv64[..., :C].reshape(B, N, h, 32).transpose(1, 2)
    v64 = kv64[..., C:].reshape(B, N, h, 32).transpose(1, 2)
    ref = torch.softmax(k64, -2).transpose(-1, -2) @ v64
    e_new, e_sep = rel_err(ctx, ref), rel_err(ref_native, ref)
    print(f"\n[kv_project_reduce B={B} N={N}] vs fp64: one launch {e_new:.2e}, projection + kv_reduce (bf16 K, V) {e_sep:.2e}")
    assert e_new < 6e-3                                          # P is rounded to bf16 (the kernels' tensor-pipe operand)
    assert rel_err(ctx, ref_native) < 1.5e-2                     # the separate path rounds K and V to bf16 as well
    assert torch.equal(ctx, ops.kv_project_reduce(x, w, b, h))    # fixed-order merge: bit-reproducible
    if B > 1:   # a sample is split by a rule of its token count only: the same bits alone or in any batch (N-GPU sliding window == 1 GPU)
        assert torch.equal(ops.kv_project_reduce(x[1:2].contiguous(), w, b, h), ctx[1:2])
        assert torch.equal(ops.kv_project_reduce(x[:B - 1].contiguous(), w, b, h), ctx[:B - 1])
    # with the output projection's weight the merge kernel also writes W_b (ctx_project of the same context)
    wo = (torch.randn(C, C, generator=g) * 0.1).to("cuda", bf)
    ctx2, wb = ops.kv_project_reduce(x, w, b, h, w_o=wo)
    assert torch.equal(ctx2, ctx) and torch.equal(wb, ops.ctx_project(ctx, wo))


def test_kv_project_reduce_rescale_path_and_limits():
    """Keys that run away from the first tile's reference (exact rescale path) and unsupported shapes."""
    ops = _ops()
    bf = torch.bfloat16
    B, N, C, h = 2, 2048, 128, 4
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, N, C, generator=g)
    x[:, N // 2:] *= 6.0                                          # later tiles carry much larger keys
    x = x.to("cuda", bf)
    w = (torch.randn(2 * C, C, generator=g) * 0.5).to("cuda", bf)
    b = torch.zeros(2 * C).cuda()
    ctx = ops.kv_project_reduce(x, w, b, h)
    kv64 = x.double() @ w.double().t() + b.double()
    k64 = kv64[..., :C].reshape(B, N, h, 32).transpose(1, 2)
    v64 = kv64[..., C:].reshape(B, N, h, 32).transpose(1, 2)
    assert float(k64[:, :, N // 2:].max() - k64[:, :, :128].max()) > 64 * math.log(2.0)      # the keys do run away from tile 0
    ref = torch.softmax(k64, -2).transpose(-1, -2) @ v64
    assert torch.isfinite(ctx).all() and rel_err(ctx, ref) < 1e-2
    assert not ops.kv_project_reduce_supported(C, h, 100) and not ops.kv_project_reduce_supported(256, 8, 512)
    with pytest.raises(RuntimeError):
        ops.kv_project_reduce(x[:, :100].contiguous(), w, b, h)


@pytest.mark.parametrize("C", [128, 256])
@pytest.mark.parametrize("rows", [1, 127, 128, 129, 1000, 148 * 128 * 2 + 77, 57408])
def test_ffn_fused(rows, C):
    """Fused FFN half of the encoder layer (trans_block.py:207-210) vs fp32 torch on the same bf16-rounded operands."""
    ops = _ops()
    l1, l2 = torch.nn.Linear(C, 2 * C), torch.nn.Linear(2 * C, C)
    with torch.no_grad():
        l1.weight.copy_(q_(l1.weight * 2, torch.bfloat16))
        l2.weight.copy_(q_(l2.weight * 2, torch.bfloat16))
    x = q_(rnd((rows, C), 90 + rows % 7, 1.5), torch.bfloat16)
    g, b = 1 + 0.1 * rnd((C,), 91), 0.1 * rnd((C,), 92)
    h = F.gelu(F.linear(x, l1.weight.detach(), l1.bias.detach()))
    ref = F.layer_norm(x + F.linear(q_(h, torch.bfloat16), l2.weight.detach(), l2.bias.detach()), (C,), g, b, eps=1e-6)
    cu = lambda t: t.detach().to("cuda")
    xd = x.to("cuda", torch.bfloat16)
    y = ops.ffn_fused(xd, cu(l1.weight).to(torch.bfloat16), cu(l1.bias), cu(l2.weight).to(torch.bfloat16), cu(l2.bias),
                      cu(g), cu(b), 1e-6)
    assert y.shape == (rows, C)
    assert rel_err(y.float(), ref) < TOL[torch.bfloat16]
    # in place, and bit-identical to the out-of-place result
    y2 = ops.ffn_fused(xd, cu(l1.weight).to(torch.bfloat16), cu(l1.bias), cu(l2.weight).to(torch.bfloat16), cu(l2.bias),
                       cu(g), cu(b), 1e-6, out=xd)
    assert torch.equal(y2, y)


@pytest.mark.parametrize("B,N", [(1, 1), (1, 128), (2, 129), (2, 300), (3, 7176), (8, 57408 // 16)])
def test_attn_out_fused(B, N):
    """Fused Q-projection + readout + output projection + residual + LayerNorm1 (trans_block.py:50,:65,:155-166,
    :205-206) vs fp32 torch on the same bf16-rounded operands."""
    ops = _ops()
    C, h = 128, 4
    lq, lo = torch.nn.Linear(C, C), torch.nn.Linear(C, C)
    with torch.no_grad():
        lq.weight.copy_(q_(lq.weight * 3, torch.bfloat16))
        lo.weight.copy_(q_(lo.weight * 3, torch.bfloat16))
    x = q_(rnd((B, N, C), 70 + N % 5, 1.5), torch.bfloat16)
    ctx = rnd((B, h, 32, 32), 71, 0.5)
    g, b = 1 + 0.1 * rnd((C,), 72), 0.1 * rnd((C,), 73)
    qh = F.linear(x, lq.weight.detach(), lq.bias.detach()).view(B, N, h, 32)
    pr = q_(torch.softmax(qh, -1) / math.sqrt(32.0), torch.bfloat16)
    att = torch.einsum("bnhj,bhje->bnhe", pr, q_(ctx, torch.bfloat16)).reshape(B, N, C)
    ref = F.layer_norm(x + F.linear(q_(att, torch.bfloat16), lo.weight.detach(), lo.bias.detach()), (C,), g, b, eps=1e-6)
    cu = lambda t: t.detach().to("cuda")
    ctx16 = ops.ctx_pack(cu(ctx).contiguous())
    ref16 = torch.zeros(B * h * 32, 64)
    ref16[:, :32] = q_(ctx, torch.bfloat16).permute(0, 1, 3, 2).reshape(B * h * 32, 32)
    assert torch.equal(ctx16.float().cpu(), ref16)
    y = ops.attn_out_fused(x.to("cuda", torch.bfloat16), cu(lq.weight).to(torch.bfloat16), cu(lq.bias), ctx16,
                           cu(lo.weight).to(torch.bfloat16), cu(lo.bias), cu(g), cu(b), h)
    assert y.shape == (B, N, C)
    assert rel_err(y.float(), ref) < TOL[torch.bfloat16]
    # the readout folded into the output projection (W_b = blockdiag(ctx_b) Wo^T): two chained GEMMs per tile
    wb = ops.ctx_project(cu(ctx).contiguous(), cu(lo.weight).to(torch.bfloat16))
    yw = ops.attn_out_fused_w(x.to("cuda", torch.bfloat16), cu(lq.weight).to(torch.bfloat16), cu(lq.bias), wb, cu(lo.bias),
                              cu(g), cu(b), h)
    ref64 = F.layer_norm(x.double() + F.linear(torch.einsum("bnhj,bhje->bnhe", pr.double(), ctx.double()).reshape(B, N, C),
                                                lo.weight.detach().double(), lo.bias.detach().double()), (C,), g.double(), b.double(), eps=1e-6)
    e_w, e_3 = rel_err(yw.float().cpu().double(), ref64), rel_err(y.float().cpu().double(), ref64)
    print(f"\n[attn_out B={B} N={N}] vs fp64: three-GEMM kernel {e_3:.2e}, folded readout {e_w:.2e}")
    assert yw.shape == (B, N, C) and e_w < TOL[torch.bfloat16] and e_w <= 1.5 * e_3 + 1e-3
    assert torch.equal(yw, ops.attn_out_fused_w(x.to("cuda", torch.bfloat16), cu(lq.weight).to(torch.bfloat16), cu(lq.bias), wb,
                                                cu(lo.bias), cu(g), cu(b), h))
