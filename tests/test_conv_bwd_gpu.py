"""GPU: backward of nn.Conv3d on the bf16 path (ltu_conv3d_wgrad on the tensor pipe; input gradient through the forward
kernels with the reversed, transposed filter) against fp64 autograd of F.conv3d on the same bf16-rounded operands.
SURVEY 8f-1."""
import pytest
import torch
import torch.nn.functional as F

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu

CASES = [
    # cin, cout, k, stride, (H, W, D), B
    (16, 16, 3, (1, 1, 1), (6, 5, 9), 2),
    (8, 16, 3, (1, 1, 1), (8, 6, 10), 1),          # stem (4 input channels padded to 8)
    (16, 32, 3, (2, 2, 1), (8, 6, 7), 2),          # DownBlock conv2, stride (2,2,1)
    (32, 64, 3, (2, 2, 2), (6, 8, 6), 2),
    (64, 64, 3, (1, 1, 1), (5, 4, 6), 3),
    (128, 128, 3, (1, 1, 1), (3, 4, 4), 2),
    (256, 128, 3, (1, 1, 1), (3, 3, 4), 1),        # several 64x64 channel blocks
    (32, 16, 1, (1, 1, 1), (5, 6, 7), 2),          # gate W_g 1x1x1
    (64, 24, 3, (1, 1, 1), (9, 7, 12), 2),         # Cout not a multiple of 16, ragged voxel tiles
]


@pytest.mark.parametrize("cin,cout,k,stride,shape,B", CASES)
def test_conv3d_backward_matches_autograd(cin, cout, k, stride, shape, B):
    from lintransunet_b200.backward import conv3d_backward
    H, W, D = shape
    torch.manual_seed(cin * 7 + cout + k)
    conv = torch.nn.Conv3d(cin, cout, k, stride=stride, padding=k // 2).cuda()
    with torch.no_grad():
        conv.weight.copy_(conv.weight.to(torch.bfloat16).float())                 # bf16-representable filter
    x = torch.randn(B, H, W, D, cin, device="cuda").to(torch.bfloat16)
    xd = x.double().permute(0, 4, 1, 2, 3).clone().requires_grad_(True)
    wd = conv.weight.detach().double().clone().requires_grad_(True)
    bd = conv.bias.detach().double().clone().requires_grad_(True)
    yd = F.conv3d(xd, wd, bd, stride=stride, padding=k // 2)
    dy = torch.randn(B, yd.shape[2], yd.shape[3], yd.shape[4], cout, device="cuda").to(torch.bfloat16)
    yd.backward(dy.double().permute(0, 4, 1, 2, 3))
    dx, dw, db = conv3d_backward(x, dy, conv)
    assert dw.shape == conv.weight.shape and dw.dtype == torch.float32
    e_w, e_b = rel_err(dw, wd.grad), rel_err(db, bd.grad)
    e_x = rel_err(dx, xd.grad.permute(0, 2, 3, 4, 1))
    print(f"\n[conv bwd {cin}->{cout} k{k} s{stride}] dW {e_w:.2e} db {e_b:.2e} dx {e_x:.2e}")
    assert e_w <= 1e-4 and e_b <= 1e-5                     # exact bf16 products, fp32 accumulation
    assert e_x <= 1e-2                                      # dx is stored in bf16
    dw2 = conv3d_backward(x, dy, conv, need_dx=False)[1]
    assert torch.equal(dw, dw2)


@pytest.mark.parametrize("cin,cout,shape", [(128, 32, (5, 4, 6)), (256, 64, (3, 4, 4))])
def test_upsampled_conv3d_backward_matches_autograd(cin, cout, shape):
    """UpEmbedBlock: nn.Upsample(nearest, x2) -> Conv3d (model/Unet_3Dblock.py:419-429)."""
    from lintransunet_b200.backward import conv3d_backward
    H, W, D = shape
    torch.manual_seed(cin + cout)
    conv = torch.nn.Conv3d(cin, cout, 3, padding=1).cuda()
    with torch.no_grad():
        conv.weight.copy_(conv.weight.to(torch.bfloat16).float())
    x = torch.randn(2, H, W, D, cin, device="cuda").to(torch.bfloat16)
    xd = x.double().permute(0, 4, 1, 2, 3).clone().requires_grad_(True)
    wd = conv.weight.detach().double().clone().requires_grad_(True)
    yd = F.conv3d(F.interpolate(xd, scale_factor=2, mode="nearest"), wd, None, padding=1)
    dy = torch.randn(2, 2 * H, 2 * W, 2 * D, cout, device="cuda").to(torch.bfloat16)
    yd.backward(dy.double().permute(0, 4, 1, 2, 3))
    dx, dw, _ = conv3d_backward(x, dy, conv, up2=True)
    e_w, e_x = rel_err(dw, wd.grad), rel_err(dx, xd.grad.permute(0, 2, 3, 4, 1))
    print(f"\n[up2 conv bwd {cin}->{cout}] dW {e_w:.2e} dx {e_x:.2e}")
    assert e_w <= 1e-4 and e_x <= 2e-2                      # dx: bf16 full-resolution gradient, then 8-term bf16 sums


def _stored(cd: torch.Tensor, raw_cl: torch.Tensor) -> torch.Tensor:
    """Straight-through substitution: the reference continues from the convolution output AS THE FORWARD STORED IT (bf16),
    gradients flow to the fp64 convolution unchanged.  Without it the comparison measures how many LeakyReLU signs of
    near-zero activations a bf16 rounding flips (dozens per layer, each worth |dy| in the max norm), not the kernels."""
    return cd + (raw_cl.double().permute(0, 4, 1, 2, 3) - cd).detach()


@pytest.mark.parametrize("cin,cout,stride,shape,residual", [(16, 16, (1, 1, 1), (12, 10, 8), True), (32, 64, (2, 2, 2), (8, 8, 8), False),
                                                            (128, 256, (2, 2, 2), (4, 4, 8), False), (128, 128, (1, 1, 1), (4, 4, 8), True)])
def test_conv_instnorm_lrelu_block_backward(cin, cout, stride, shape, residual):
    """One Conv3d -> InstanceNorm3d -> LeakyReLU (+ residual) stage of DownBlock (model/Unet_3Dblock.py:325-336): input and
    parameter gradients against fp64 autograd on the same bf16 input and the same stored convolution output."""
    from lintransunet_b200.backward import conv_in_act_backward, conv_in_act_train
    H, W, D = shape
    torch.manual_seed(cin + cout)
    conv = torch.nn.Conv3d(cin, cout, 3, stride=stride, padding=1).cuda()
    with torch.no_grad():
        conv.weight.copy_(conv.weight.to(torch.bfloat16).float())
    x = torch.randn(2, H, W, D, cin, device="cuda").to(torch.bfloat16)
    y, saved = conv_in_act_train(x, conv, residual=x if residual else None)
    xd = x.double().permute(0, 4, 1, 2, 3).clone().requires_grad_(True)
    wd = conv.weight.detach().double().clone().requires_grad_(True)
    bd = conv.bias.detach().double().clone().requires_grad_(True)
    cd = F.conv3d(xd, wd, bd, stride=stride, padding=1)
    assert rel_err(saved["raw"], cd.detach().permute(0, 2, 3, 4, 1)) <= 1e-2
    yd = F.leaky_relu(F.instance_norm(_stored(cd, saved["raw"]), eps=1e-5), 0.01)
    if residual:
        yd = yd + xd
    assert rel_err(y, yd.detach().permute(0, 2, 3, 4, 1)) <= 1e-2
    dy = torch.randn(y.shape, device="cuda").to(torch.bfloat16)
    yd.backward(dy.double().permute(0, 4, 1, 2, 3))
    dx, dw, db = conv_in_act_backward(dy, saved)
    if residual:
        dx = dx + dy
    e_x, e_w = rel_err(dx, xd.grad.permute(0, 2, 3, 4, 1)), rel_err(dw, wd.grad)
    print(f"\n[conv+IN+LReLU bwd {cin}->{cout} s{stride}] dx {e_x:.2e} dW {e_w:.2e}")
    assert e_x <= 1.5e-2 and e_w <= 1e-2          # the gradient of the raw output is stored in bf16 before it is contracted


def test_encoder_backward_matches_autograd_on_the_stored_activations():
    """All 18 parameter gradients of Encoder (stem + 4 DownBlocks, model/Unet_3Dblock.py:596-607: restated here like
    oracle.encoder_forward, plus the straight-through substitution of every stored convolution output)."""
    from oracle import ltu_oracle as O
    from lintransunet_b200.backward import encoder_backward, encoder_train
    from lintransunet_b200.unet import Encoder
    torch.manual_seed(5)
    cfg = O.UnetConfig(dim_output=2)
    enc = Encoder(list(cfg.num_layers), 1).cuda()
    with torch.no_grad():
        for p_ in enc.parameters():
            p_.copy_(p_.to(torch.bfloat16).float())
    x = torch.randn(2, 1, 64, 64, 16, device="cuda")
    bottle, skips, saved = encoder_train(x, enc)
    sd = {k: v.detach().double().clone().requires_grad_(True) for k, v in enc.state_dict().items()}
    cia = lambda cd, sv, res=None: O.lrelu(O.inorm(_stored(cd, sv["raw"]))) + (0 if res is None else res)
    a = O.space_to_depth(x.to(torch.bfloat16).double(), 2)                      # the kernels read the bf16 input
    a = cia(F.conv3d(a, sd["input_block.weight"], sd["input_block.bias"], padding=1), saved["stem"])
    skips_d = []
    for i, (sv1, sv2) in enumerate(saved["blocks"]):
        p_ = f"block_list.{i}"
        s_ = cia(F.conv3d(a, sd[p_ + ".conv1.weight"], sd[p_ + ".conv1.bias"], padding=1), sv1, a)
        skips_d.append(s_)
        a = cia(F.conv3d(s_, sd[p_ + ".conv2.weight"], sd[p_ + ".conv2.bias"], stride=(2, 2, i % 2 + 1), padding=1), sv2)
    to_cl = lambda t: t.permute(0, 2, 3, 4, 1)
    assert rel_err(bottle, to_cl(a.detach())) <= 2e-2
    g = torch.Generator(device="cuda").manual_seed(9)
    d_bottle = torch.randn(bottle.shape, device="cuda", generator=g).to(torch.bfloat16)
    d_skips = [torch.randn(s.shape, device="cuda", generator=g).to(torch.bfloat16) for s in skips]
    loss = (a * d_bottle.double().permute(0, 4, 1, 2, 3)).sum()
    for s_d, ds in zip(skips_d, d_skips):
        loss = loss + (s_d * ds.double().permute(0, 4, 1, 2, 3)).sum()
    loss.backward()
    grads = encoder_backward(d_bottle, d_skips, saved)
    assert sorted(grads) == sorted(sd)
    worst, worst_name = 0.0, ""
    for name, gr in sorted(grads.items()):
        ref = sd[name].grad
        assert gr.shape == ref.shape, name
        if name.endswith(".bias"):
            # a bias in front of an InstanceNorm has a mathematically zero gradient: noise on both sides
            assert float(gr.abs().max()) <= 5e-2 * float(grads[name[:-4] + "weight"].abs().max()), name
            continue
        e = rel_err(gr, ref)
        if e > worst:
            worst, worst_name = e, name
    print(f"\n[encoder bwd bf16] worst weight-gradient error {worst:.2e} ({worst_name})")
    assert worst <= 5e-2                           # nine stages of bf16 gradients


def test_embed_attention_block_backward():
    """EmbedAttention3DBlock (the inside of a ROI bridge, model/Unet_3Dblock.py:469-501): down_embed conv -> 8 encoder
    layers + positional conv -> upsample + up_embed conv.  Reference: fp64 autograd through the oracle's transformer
    stack, continuing from the two convolution outputs as stored."""
    from oracle import ltu_oracle as O
    from lintransunet_b200.backward import embed_block_backward, embed_block_train
    from lintransunet_b200.unet import EmbedAttention3DBlock
    torch.manual_seed(21)
    in_dim, C, nhead, n_layers = 32, 128, 4, 8
    blk = EmbedAttention3DBlock(in_dim, C, nhead, n_layers).cuda()
    with torch.no_grad():
        for p_ in blk.parameters():
            p_.copy_(p_.to(torch.bfloat16).float())
    x = torch.randn(2, 16, 12, 8, in_dim, device="cuda").to(torch.bfloat16)
    y, saved = embed_block_train(x, blk)
    sd = {f"T.{k}": v.detach().double().clone().requires_grad_(True) for k, v in blk.state_dict().items()}
    xd = x.double().permute(0, 4, 1, 2, 3).clone().requires_grad_(True)
    act = lambda cd, sv: O.lrelu(O.inorm(_stored(cd, sv["raw"])))
    t = act(F.conv3d(xd, sd["T.down_embed.module_list.0.0.weight"], sd["T.down_embed.module_list.0.0.bias"], stride=2, padding=1),
            saved["down"])
    t = O.transformer_stack(t, sd, "T", nhead, sd["T.pos_encoder.proj.weight"], sd["T.pos_encoder.proj.bias"], n_layers)
    t = F.interpolate(t, scale_factor=2, mode="nearest")
    yd = act(F.conv3d(t, sd["T.up_embed.module_list.0.1.weight"], sd["T.up_embed.module_list.0.1.bias"], padding=1), saved["up"])
    assert y.shape == x.shape and rel_err(y, yd.detach().permute(0, 2, 3, 4, 1)) <= 2e-2
    dy = torch.randn(y.shape, device="cuda").to(torch.bfloat16)
    yd.backward(dy.double().permute(0, 4, 1, 2, 3))
    dx, grads = embed_block_backward(dy, saved)
    assert sorted(grads) == sorted(k[2:] for k in sd)
    e_x = rel_err(dx, xd.grad.permute(0, 2, 3, 4, 1))
    worst, worst_name = 0.0, ""
    for name, gr in grads.items():
        ref = sd["T." + name].grad
        assert gr.shape == ref.shape, name
        if name.endswith("self_attn.linears.1.bias") or (name.endswith(".bias") and "embed" in name):
            continue                                        # mathematically zero gradients (softmax over tokens / InstanceNorm)
        e = rel_err(gr, ref)
        if e > worst:
            worst, worst_name = e, name
    print(f"\n[embed block bwd bf16] dx rel err {e_x:.2e}, worst parameter gradient {worst:.2e} ({worst_name})")
    assert e_x <= 8e-2 and worst <= 1.2e-1          # measured 1.2e-2 / 5.7e-2; bit-reproducible kernels, fixed seeds
