"""Training-mode forward and backward of the model's building blocks on the sm_100a kernels.

Every operator of ``MaskTransUnet`` has a gradient kernel (include/ltu_b200.h: ``ltu_attn_bwd``, ``ltu_add_layernorm_bwd``,
``ltu_gelu_bwd``, ``ltu_posenc_wgrad``, ``ltu_instnorm_bwd``, ``ltu_conv3d_wgrad`` + the forward convolution kernels for the
input gradient, ``ltu_upsample_trilinear_bwd``, ``ltu_roi_resample_bwd``, ``ltu_gate_bwd``, ``ltu_mask_softmax_bwd``,
``ltu_head_d2s_softmax_bwd``); this module composes them, block by block, each pair ``*_train`` / ``*_backward`` with a parity
test against fp64 autograd through the oracle (tests/test_encoder_layer_bwd_gpu.py, test_conv_bwd_gpu.py, test_unet_bwd_gpu.py):

* ``encoder_layer_*``      one SelfAttentionLayer (model/trans_block.py:169-211), fp32 or bf16
* ``transformer_stack_*``  the 8-layer stack with its positional conv (PosAttention3DBlock / EmbedAttention3DBlock)
* ``conv3d_backward``, ``conv_in_act_*``   Conv3d (strided / upsampled) and Conv -> InstanceNorm -> LeakyReLU (+ residual), bf16
* ``encoder_*``            the whole Encoder (stem + 4 DownBlocks)
* ``embed_block_*``, ``roi_bridge_*``      the inside of a ROI bridge and the bridge itself (fisheye resample both ways)
* ``upblock_*``, ``gate_*``                the decoder's UpBlock and attention gate

``head_conv_*`` (the 2/3/12-channel head convolutions, output gradient padded to 8 channels), ``decoder_*`` (the ROIDecoder
loop) and ``model_loss_and_gradients`` (encoder + decoder + ``lintransunet_b200.losses``) complete the chain: one training
step's loss terms agree with the unmodified reference to 1e-3 and every parameter-gradient norm to 4.6e-2 on B200
(tests/test_train_step_gpu.py, bf16 activations against the reference's fp32 run), and the host logic reproduces all 600
gradients of oracle autograd to 3e-8 when every kernel wrapper is replaced by an fp64 torch stand-in
(tests/test_backward_composition_cpu.py).  The ``autograd.Function`` wiring (``unet._NativeTrainFunction``) calls the same
functions; it is verified with the stand-ins only and therefore still opt-in (``model.native_backward``).  Training-mode dropout (the reference's
default p = 0.3) is a Philox mask re-generated in the backward (``DropoutState``, ``ltu_dropout``); the parity tests run
dropout-free, the dropout path has its own statistical and gradient-consistency tests (tests/test_dropout_gpu.py).

The ``nn.Linear`` layers and their weight gradients are plain cuBLAS GEMMs (``F.linear`` / ``torch.mm`` / ``torch.addmm``), as
in the forward; small tensor glue (concatenation, residual adds, bias-gradient row sums) uses torch ops.  Packed / transposed
weights are cached per parameter object and version (``_memo``): an optimizer step re-packs once, nothing is built on the CPU.  Parameter gradients are returned in fp32 under the reference's parameter names.
"""
from __future__ import annotations

import weakref
from types import SimpleNamespace
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

from . import ops

_ACT = torch.bfloat16          # activation storage type of the convolutional blocks (the conv weight gradient is bf16 only)

__all__ = ["encoder_layer_train", "encoder_layer_backward", "transformer_stack_train", "transformer_stack_backward",
           "conv3d_backward", "conv_in_act_train", "conv_in_act_backward", "encoder_train", "encoder_backward",
           "embed_block_train", "embed_block_backward", "upblock_train", "upblock_backward", "gate_train", "gate_backward", "roi_bridge_train", "roi_bridge_backward",
           "head_conv_train", "head_conv_backward", "decoder_train", "decoder_backward", "model_loss_and_gradients"]


# Derived (packed / transposed / bf16) weights of the training path, keyed on the parameter OBJECTS and their versions: an
# optimizer step updates a parameter in place (same object, new version -> rebuilt once), a new model is a new object.
# One slot per (tag, parameters): stale generations are overwritten, not accumulated.
_PACKS: Dict[tuple, tuple] = {}


def _memo(tag, tensors, build):
    key = (tag,) + tuple(id(t) for t in tensors)
    ver = tuple(t._version for t in tensors) + tuple(t.data_ptr() for t in tensors)
    hit = _PACKS.get(key)
    if hit is not None and hit[0] == ver and all(r() is t for r, t in zip(hit[1], tensors)):
        return hit[2]
    val = build()
    if len(_PACKS) > 4096:
        _PACKS.clear()
    _PACKS[key] = (ver, tuple(weakref.ref(t) for t in tensors), val)
    return val


def _conv_pack(conv, **kw):
    """unet._ConvW of an nn.Conv3d container, cached across training steps."""
    from .unet import _ConvW
    ts = [conv.weight] + ([conv.bias] if conv.bias is not None else [])
    return _memo(("conv",) + tuple(sorted(kw.items())), ts, lambda: _ConvW(conv, True, **kw))


class DropoutState:
    """The dropout stream of ONE training forward (p = the model's `dropout`, 0.3 in the reference): every site draws from
    Philox4x32-10 keyed by the CUDA generator's seed at consecutive counter offsets (ops.dropout), and remembers its offset
    in the saved state so that the backward re-applies the identical mask to the gradient.  `finish()` advances torch's
    CUDA generator past the counters used, so the next forward (and torch's own random ops) draw fresh numbers and
    `torch.manual_seed` makes a training run reproducible.  The CPU generator is never touched."""

    def __init__(self, p: float = 0.0, device: torch.device = None, seed: int = None, offset: int = 0):
        self.p = float(p or 0.0)
        self._gen = None
        self.seed, self.offset = 0, int(offset)
        if self.p > 0.0:
            if not 0.0 < self.p < 1.0:
                raise ValueError(f"dropout must be in [0, 1), got {p}")
            if seed is not None:
                self.seed = int(seed)
            else:
                idx = device.index if device.index is not None else torch.cuda.current_device()
                self._gen = torch.cuda.default_generators[idx]
                self.seed, self.offset = self._gen.initial_seed(), self._gen.get_offset()

    def apply(self, x: torch.Tensor, channelwise: bool = False):
        """In-place dropout of a freshly produced activation; returns (x, token for `backward`)."""
        if self.p == 0.0:
            return x, None
        tok = (self.offset, channelwise)
        ops.dropout(x, self.p, self.seed, self.offset, channelwise, inplace=True)
        self.offset += ops.dropout_counters(x, channelwise)
        return x, tok

    def backward(self, dy: torch.Tensor, tok):
        if tok is None:
            return dy
        return ops.dropout(dy.contiguous(), self.p, self.seed, tok[0], tok[1], inplace=False)

    def finish(self) -> None:
        if self._gen is not None:
            self._gen.set_offset((self.offset + 3) // 4 * 4)


_NO_DROP = DropoutState(0.0)


def _params(layer, dtype: torch.dtype) -> Dict[str, torch.Tensor]:
    # attribute access, not .parameters(): an nn.DataParallel replica keeps its weights as plain attributes
    mods = list(layer.self_attn.linears) + [layer.linear1, layer.linear2, layer.layer_norm1, layer.layer_norm2]
    ts = [t for m in mods for t in (m.weight, m.bias)]
    return _memo(("layer", dtype), ts, lambda: _params_build(layer, dtype))


def _params_build(layer, dtype: torch.dtype) -> Dict[str, torch.Tensor]:
    lin = layer.self_attn.linears
    c = lambda t: t.detach().to(dtype).contiguous()
    f = lambda t: t.detach().float().contiguous()
    return dict(w_qkv=c(torch.cat([lin[0].weight, lin[1].weight, lin[2].weight], 0)),
                b_qkv=c(torch.cat([lin[0].bias, lin[1].bias, lin[2].bias], 0)),
                w_o=c(lin[3].weight), b_o=c(lin[3].bias), w_1=c(layer.linear1.weight), b_1=c(layer.linear1.bias),
                w_2=c(layer.linear2.weight), b_2=c(layer.linear2.bias),
                g1=f(layer.layer_norm1.weight), be1=f(layer.layer_norm1.bias),
                g2=f(layer.layer_norm2.weight), be2=f(layer.layer_norm2.bias), nhead=layer.self_attn.nhead)


@torch.no_grad()
def encoder_layer_train(t: torch.Tensor, layer, drop: DropoutState = _NO_DROP) -> Tuple[torch.Tensor, dict]:
    """SelfAttentionLayer.forward on tokens [B,N,C] (fp32 or bf16, CUDA), keeping what the backward needs.
    `layer` is the parameter container lintransunet_b200.unet.SelfAttentionLayer.  Training-mode dropout (`drop`):
    dropout1 on the attention output (:205), dropout on the GELU output (:208), dropout2 on linear2's output (:209); the
    dropout inside linear_attention (:62-63) does not reach its result in the reference and draws nothing here."""
    p = _params(layer, t.dtype)
    B, N, C = t.shape
    h = p["nhead"]
    qkv = F.linear(t, p["w_qkv"], p["b_qkv"])
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    ctx = ops.kv_reduce(k, v, h)
    att = ops.q_readout(q, ctx, h)
    o, tok_o = drop.apply(F.linear(att, p["w_o"], p["b_o"]))
    t1 = ops.add_layernorm(t, o, p["g1"], p["be1"], 1e-6)
    f1 = F.linear(t1, p["w_1"], p["b_1"])
    fa, tok_fa = drop.apply(ops.gelu(f1))
    f2, tok_f2 = drop.apply(F.linear(fa, p["w_2"], p["b_2"]))
    t2 = ops.add_layernorm(t1, f2, p["g2"], p["be2"], 1e-6)
    return t2, dict(p=p, t=t, qkv=qkv, ctx=ctx, att=att, o=o, t1=t1, f1=f1, fa=fa, f2=f2, drop=drop,
                    toks=(tok_o, tok_fa, tok_f2))


def _wgrad(dy2: torch.Tensor, x2: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """y = x W^T + b  ->  dW = dy^T x, db = sum_rows dy (fp32)."""
    return torch.mm(dy2.t(), x2).float(), torch.sum(dy2, 0, dtype=torch.float32)


@torch.no_grad()
def encoder_layer_backward(dout: torch.Tensor, saved: dict) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """Gradient of the layer input and of the 16 parameters, keyed like the reference's state_dict entries
    relative to the layer (``self_attn.linears.0.weight`` ... ``layer_norm2.bias``)."""
    p, t, t1 = saved["p"], saved["t"], saved["t1"]
    B, N, C = t.shape
    h = p["nhead"]
    rows = B * N
    g: Dict[str, torch.Tensor] = {}
    two = lambda a: a.reshape(rows, a.shape[-1])
    drop = saved.get("drop", _NO_DROP)
    tok_o, tok_fa, tok_f2 = saved.get("toks", (None, None, None))
    # y = LN2(t1 + dropout2(f2));  saved f2 / fa / o are the tensors AFTER their dropout
    dz2, g["layer_norm2.weight"], g["layer_norm2.bias"] = ops.add_layernorm_bwd(t1, saved["f2"], dout.contiguous(), p["g2"], 1e-6)
    df2 = drop.backward(dz2, tok_f2)
    g["linear2.weight"], g["linear2.bias"] = _wgrad(two(df2), two(saved["fa"]))
    dfa = drop.backward(torch.mm(two(df2), p["w_2"]).reshape(B, N, 2 * C), tok_fa)
    df1 = ops.gelu_bwd(saved["f1"], dfa)
    g["linear1.weight"], g["linear1.bias"] = _wgrad(two(df1), two(t1))
    dt1 = torch.addmm(two(dz2), two(df1), p["w_1"]).reshape(B, N, C)            # residual + through linear1
    # t1 = LN1(t + dropout1(o))
    dz1, g["layer_norm1.weight"], g["layer_norm1.bias"] = ops.add_layernorm_bwd(t, saved["o"], dt1, p["g1"], 1e-6)
    do = drop.backward(dz1, tok_o)
    g["self_attn.linears.3.weight"], g["self_attn.linears.3.bias"] = _wgrad(two(do), two(saved["att"]))
    datt = torch.mm(two(do), p["w_o"]).reshape(B, N, C)
    qkv = saved["qkv"]
    dqkv = ops.linear_attention_bwd(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], saved["ctx"], datt, h)
    dw, db = _wgrad(two(dqkv), two(t))
    for i in range(3):
        g[f"self_attn.linears.{i}.weight"] = dw[i * C:(i + 1) * C]
        g[f"self_attn.linears.{i}.bias"] = db[i * C:(i + 1) * C]
    dt = torch.addmm(two(dz1), two(dqkv), p["w_qkv"]).reshape(B, N, C)          # residual + through the QKV projection
    return dt, g


@torch.no_grad()
def transformer_stack_train(x: torch.Tensor, layers, pos_encoder, drop: DropoutState = _NO_DROP) -> Tuple[torch.Tensor, dict]:
    """The 8-layer stack of PosAttention3DBlock / EmbedAttention3DBlock (model/Unet_3Dblock.py:265-270, :484-490) on
    a channels-last volume [B,H,W,D,C]: positional depthwise conv once, after layer 0.  `layers` = the ModuleList of
    SelfAttentionLayer containers, `pos_encoder` = the Conv3dPosEmbedding container that is actually used."""
    from .unet import _pos_w
    B, H, W, D, C = x.shape
    w27, pb = _pos_w(pos_encoder)
    t = x.reshape(B, H * W * D, C)
    saved_layers, pos_in, tok_pos = [], None, None
    for i, layer in enumerate(layers):
        t, sv = encoder_layer_train(t, layer, drop)
        saved_layers.append(sv)
        if i == 0:
            pos_in = t.reshape(B, H, W, D, C)
            t, tok_pos = drop.apply(ops.posenc_dwconv3(pos_in, w27, pb), channelwise=True)     # nn.Dropout3d, trans_block.py:96
            t = t.reshape(B, H * W * D, C)
    return t.reshape(B, H, W, D, C), dict(layers=saved_layers, pos_in=pos_in, w27=w27, shape=(B, H, W, D, C), drop=drop,
                                          tok_pos=tok_pos)


@torch.no_grad()
def transformer_stack_backward(dout: torch.Tensor, saved: dict) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """Gradients of the stack input [B,H,W,D,C] and of its parameters, keyed ``layers.<i>.<name>`` and
    ``pos.proj.weight`` [C,1,3,3,3] / ``pos.proj.bias`` (the reference's parameter layout: kernel axes (kd,kh,kw))."""
    B, H, W, D, C = saved["shape"]
    grads: Dict[str, torch.Tensor] = {}
    dt = dout.reshape(B, H * W * D, C).contiguous()
    for i in range(len(saved["layers"]) - 1, -1, -1):
        if i == 0:
            dt = saved.get("drop", _NO_DROP).backward(dt.reshape(B, H, W, D, C), saved.get("tok_pos"))
            dvol, dw27, db = ops.posenc_dwconv3_bwd(saved["pos_in"], dt.reshape(B, H, W, D, C), saved["w27"])
            # packed [kh,kw,kd][c] = weight[c,0,kd,kh,kw] (unet._pos_w)
            grads["pos.proj.weight"] = dw27.reshape(3, 3, 3, C).permute(3, 2, 0, 1).unsqueeze(1).contiguous()
            grads["pos.proj.bias"] = db
            dt = dvol.reshape(B, H * W * D, C)
        dt, g = encoder_layer_backward(dt, saved["layers"][i])
        for k, v in g.items():
            grads[f"layers.{i}.{k}"] = v
    return dt.reshape(B, H, W, D, C), grads


@torch.no_grad()
def conv3d_backward(x: torch.Tensor, dy: torch.Tensor, conv: torch.nn.Conv3d, need_dx: bool = True, up2: bool = False):
    """Backward of a 3x3x3 (pad 1) or 1x1x1 (pad 0) nn.Conv3d on channels-last bf16 tensors: returns (dx | None, dW in
    the parameter layout [Cout,Cin,k,k,k] fp32, dbias fp32).  `up2`: the convolution read nn.Upsample(nearest, x2)(x)
    (UpEmbedBlock).  dW is ltu_conv3d_wgrad.  dx is the FORWARD convolution kernel applied to the output gradient with
    the filter reversed and its channel axes swapped -- after zero insertion for a strided convolution, followed by
    2x2x2 block sums for up2."""
    from .unet import _ConvW
    k = conv.weight.shape[2]
    stride = tuple(conv.stride)
    pad = k // 2
    cout, cin = conv.weight.shape[0], conv.weight.shape[1]
    cin_x = x.shape[-1]                                    # > cin for the zero-padded stem input
    dw = ops.conv3d_wgrad(x, dy, k, stride, pad, up2=up2).reshape(k, k, k, cout, cin_x).permute(3, 4, 0, 1, 2).contiguous()
    db = torch.sum(dy.reshape(-1, cout), 0, dtype=torch.float32)
    dx = None
    if need_dx:
        # transposed filter built on the device from the weight itself: no module construction, no RNG, no H2D copy
        cw = _memo(("convT",), [conv.weight], lambda: _ConvW(SimpleNamespace(
            weight=conv.weight.detach().flip(2, 3, 4).transpose(0, 1).contiguous(), bias=None, stride=(1, 1, 1)), True))
        e = 2 if up2 else 1
        full = (e * x.shape[1], e * x.shape[2], e * x.shape[3])          # extent the convolution actually read
        z = dy if stride == (1, 1, 1) else ops.zero_insert(dy, full, stride)
        dx = ops.conv3d(z, cw.w, None, cw.cout, k, pad=pad, w_tc=cw.w_tc)[0]
        if up2:
            dx = ops.sumpool2(dx)
    return dx, dw, db


# ----------------------------------------------------------------------------- conv blocks (bf16 path)
@torch.no_grad()
def conv_in_act_train(x: torch.Tensor, conv: torch.nn.Conv3d, residual=None, cin_pad: int = 0, up2: bool = False,
                      drop: DropoutState = _NO_DROP):
    """Conv3d -> InstanceNorm3d -> LeakyReLU (+ residual) keeping the raw convolution output and the statistics
    (DownBlock / UpBlock / Encoder stem / embed blocks, model/Unet_3Dblock.py:325-336,:547-554,:596-600,:373-382,
    :419-429).  `up2`: nn.Upsample(nearest, x2) in front of the convolution (UpEmbedBlock).  `drop`: the nn.Dropout that
    follows the activation in DownBlock.conv2 (:339), the embed blocks (:382,:429) and UpBlock.conv2 (:556).  bf16,
    channels-last."""
    cw = _conv_pack(conv, cin_pad=cin_pad, fold_up2=up2)
    stride = tuple(conv.stride)
    raw, partials, _ = ops.conv3d(x, cw.w, cw.b, cw.cout, cw.k, stride=stride, pad=cw.k // 2, up2=up2, want_stats=True,
                                  w_tc=cw.w_tc, w_tc_fold=cw.w_tc_fold)
    V = raw.shape[1] * raw.shape[2] * raw.shape[3]
    stats = ops.instnorm_finalize(partials, V)
    y, tok = drop.apply(ops.instnorm_apply(raw, stats, ops.ACT_LRELU, residual=residual, inplace=False))
    return y, dict(x=x, raw=raw, stats=stats, conv=conv, residual=residual is not None, up2=up2, drop=drop, tok=tok)


@torch.no_grad()
def conv_in_act_backward(dy: torch.Tensor, saved: dict, need_dx: bool = True):
    """Returns (dx | None, dW, dbias); a residual added after the activation receives dy itself (the caller adds it)."""
    dy = saved.get("drop", _NO_DROP).backward(dy, saved.get("tok"))
    draw = ops.instnorm_bwd(saved["raw"], saved["stats"], dy.contiguous(), ops.ACT_LRELU)
    return conv3d_backward(saved["x"], draw, saved["conv"], need_dx=need_dx, up2=saved.get("up2", False))


@torch.no_grad()
def encoder_train(x: torch.Tensor, enc, drop: DropoutState = _NO_DROP):
    """Encoder.forward (model/Unet_3Dblock.py:596-607) on the bf16 path with everything the backward needs.
    x fp32 [B,1,H,W,D]; `enc` = the lintransunet_b200.unet.Encoder container.  Returns (bottleneck, skips, saved)."""
    a = ops.s2d_input(x.contiguous().float(), _ACT, cpad=8)                    # 4 channels + 4 zero channels
    a, sv_stem = conv_in_act_train(a, enc.input_block, cin_pad=8)
    blocks, skips = [], []
    for blk in enc.block_list:
        s, sv1 = conv_in_act_train(a, blk.conv1, residual=a)                  # DownBlock :327-331
        skips.append(s)
        a, sv2 = conv_in_act_train(s, blk.conv2, drop=drop)                   # :335-339 (strided; dropout on x, not on the skip)
        blocks.append((sv1, sv2))
    return a, skips, dict(stem=sv_stem, blocks=blocks)


@torch.no_grad()
def encoder_backward(d_bottle: torch.Tensor, d_skips, saved: dict) -> Dict[str, torch.Tensor]:
    """Parameter gradients of the Encoder, keyed like its state_dict (``input_block.weight`` ...,
    ``block_list.<i>.conv{1,2}.{weight,bias}``), given the gradients of the bottleneck and of the four skips
    (None = no gradient flows into that skip)."""
    grads: Dict[str, torch.Tensor] = {}
    da = d_bottle
    for i in range(len(saved["blocks"]) - 1, -1, -1):
        sv1, sv2 = saved["blocks"][i]
        ds, grads[f"block_list.{i}.conv2.weight"], grads[f"block_list.{i}.conv2.bias"] = conv_in_act_backward(da, sv2)
        if d_skips is not None and d_skips[i] is not None:
            ds = ds + d_skips[i]
        dx1, grads[f"block_list.{i}.conv1.weight"], grads[f"block_list.{i}.conv1.bias"] = conv_in_act_backward(ds, sv1)
        da = dx1 + ds                                                          # s = act(norm(conv1(a))) + a
    _, dw, grads["input_block.bias"] = conv_in_act_backward(da, saved["stem"], need_dx=False)
    grads["input_block.weight"] = dw[:, :4].contiguous()                       # drop the four zero-padded input channels
    return grads


@torch.no_grad()
def embed_block_train(x: torch.Tensor, blk, drop: DropoutState = _NO_DROP):
    """EmbedAttention3DBlock.forward (model/Unet_3Dblock.py:469-501), the inside of a ROI bridge: stride-2 down_embed
    conv + IN + LeakyReLU -> 8 encoder layers with the positional conv -> nearest x2 + up_embed conv + IN + LeakyReLU.
    x bf16 [B,h,w,d,in_dim]; `blk` = lintransunet_b200.unet.EmbedAttention3DBlock."""
    t, sv_down = conv_in_act_train(x, blk.down_embed.conv, drop=drop)
    t, sv_stack = transformer_stack_train(t, blk.layers, blk.pos_encoder, drop)
    y, sv_up = conv_in_act_train(t, blk.up_embed.conv, up2=True, drop=drop)
    return y, dict(down=sv_down, stack=sv_stack, up=sv_up)


@torch.no_grad()
def embed_block_backward(dy: torch.Tensor, saved: dict):
    """Gradients of the block input and of all its parameters, keyed like the block's state_dict."""
    grads: Dict[str, torch.Tensor] = {}
    dt, grads["up_embed.module_list.0.1.weight"], grads["up_embed.module_list.0.1.bias"] = conv_in_act_backward(dy, saved["up"])
    dt, g = transformer_stack_backward(dt, saved["stack"])
    for k, v in g.items():
        grads[k.replace("pos.", "pos_encoder.")] = v
    dx, grads["down_embed.module_list.0.0.weight"], grads["down_embed.module_list.0.0.bias"] = conv_in_act_backward(dt, saved["down"])
    return dx, grads


@torch.no_grad()
def upblock_train(x: torch.Tensor, skip: torch.Tensor, blk, drop: DropoutState = _NO_DROP):
    """UpBlock.forward (model/Unet_3Dblock.py:540-557): conv1 + IN + LeakyReLU, concatenation with the (gated, bridged)
    skip, conv2 + IN + LeakyReLU.  The training path materialises the concatenation (its gradient is a split)."""
    x1, sv1 = conv_in_act_train(x, blk.conv1)
    cat = torch.cat([x1, skip], -1)
    y, sv2 = conv_in_act_train(cat, blk.conv2, drop=drop)                     # dropout after the second activation (:555-556)
    return y, dict(c1=sv1, c2=sv2, split=x1.shape[-1])


@torch.no_grad()
def upblock_backward(dy: torch.Tensor, saved: dict):
    """Returns (dx, dskip, parameter gradients keyed conv{1,2}.{weight,bias})."""
    g: Dict[str, torch.Tensor] = {}
    dcat, g["conv2.weight"], g["conv2.bias"] = conv_in_act_backward(dy, saved["c2"])
    c = saved["split"]
    dx, g["conv1.weight"], g["conv1.bias"] = conv_in_act_backward(dcat[..., :c].contiguous(), saved["c1"])
    return dx, dcat[..., c:].contiguous(), g


@torch.no_grad()
def gate_train(skip: torch.Tensor, up: torch.Tensor, att):
    """SpatialAttention3DBlock + `encoded * attn` (model/Unet_3Dblock.py:217-221,:1385): skip [B,h,w,d,C],
    up [B,h,w,d,Cg] (the upsampled decoder state), `att` = lintransunet_b200.unet.SpatialAttention3DBlock."""
    wx, wg = _conv_pack(att.W_x[0]), _conv_pack(att.W_g[0])
    ga, pa, _ = ops.conv3d(skip, wx.w, wx.b, wx.cout, 1, pad=0, want_stats=True, w_tc=wx.w_tc)
    gg, pg, _ = ops.conv3d(up, wg.w, wg.b, wg.cout, 1, pad=0, want_stats=True, w_tc=wg.w_tc)
    V = skip.shape[1] * skip.shape[2] * skip.shape[3]
    sa, sg = ops.instnorm_finalize(pa, V), ops.instnorm_finalize(pg, V)
    psi_w, psi_b = _memo(("psi",), [att.psi[0].weight, att.psi[0].bias],
                         lambda: (att.psi[0].weight.detach().reshape(-1).float().contiguous(),
                                  att.psi[0].bias.detach().float().contiguous()))
    out = ops.gate_fused(ga, sa, gg, sg, psi_w, psi_b, skip)
    return out, dict(skip=skip, up=up, ga=ga, gg=gg, sa=sa, sg=sg, psi_w=psi_w, psi_b=psi_b, att=att)


@torch.no_grad()
def gate_backward(dout: torch.Tensor, saved: dict):
    """Returns (dskip, dup, gradients keyed like the block's state_dict: W_x.0.*, W_g.0.*, psi.0.*)."""
    att = saved["att"]
    g: Dict[str, torch.Tensor] = {}
    dskip, dh, dpw, dpb = ops.gate_bwd(saved["ga"], saved["sa"], saved["gg"], saved["sg"], saved["psi_w"], saved["psi_b"],
                                       saved["skip"], dout.contiguous())
    g["psi.0.weight"] = dpw.reshape(att.psi[0].weight.shape)
    g["psi.0.bias"] = dpb
    da = ops.instnorm_bwd(saved["ga"], saved["sa"], dh, ops.ACT_NONE)
    dg = ops.instnorm_bwd(saved["gg"], saved["sg"], dh, ops.ACT_NONE)
    dsk2, g["W_x.0.weight"], g["W_x.0.bias"] = conv3d_backward(saved["skip"], da, att.W_x[0])
    dup, g["W_g.0.weight"], g["W_g.0.bias"] = conv3d_backward(saved["up"], dg, att.W_g[0])
    return dskip + dsk2, dup, g


@torch.no_grad()
def roi_bridge_train(skip: torch.Tensor, fg: torch.Tensor, bridge, box: torch.Tensor = None, drop: DropoutState = _NO_DROP):
    """ROIBridge.forward (model/Unet_3Dblock.py:717-755): box from the foreground probability (not differentiated, as in
    the reference where it comes out of searchsorted), fisheye resample into the ROI, EmbedAttention3DBlock, resample
    back.  skip bf16 [B,h,w,d,C], fg fp32 [B,h,w,d]; `bridge` = lintransunet_b200.unet.ROIBridge."""
    B, h, w, d, C = skip.shape
    if box is None:
        box = ops.roi_bbox(fg, bridge.min_h_roi, bridge.min_w_roi, bridge.mask_threshold)
    geo = (bridge.h_roi_size, bridge.w_roi_size, bridge.eval_h_roi_size, bridge.eval_w_roi_size)
    roi = ops.roi_resample(skip, box, (h, w), *geo, direction=0)
    t, sv = embed_block_train(roi, bridge.transformer, drop)
    out = ops.roi_resample(t, box, (h, w), *geo, direction=1)
    return out, dict(box=box, geo=geo, hw=(h, w), block=sv)


@torch.no_grad()
def roi_bridge_backward(dout: torch.Tensor, saved: dict):
    """Returns (dskip, parameter gradients keyed like the bridge's state_dict: transformer.<...>)."""
    dt = ops.roi_resample_bwd(dout.contiguous(), saved["box"], saved["hw"], *saved["geo"], direction=1)
    droi, g = embed_block_backward(dt, saved["block"])
    dskip = ops.roi_resample_bwd(droi, saved["box"], saved["hw"], *saved["geo"], direction=0)
    return dskip, {f"transformer.{k}": v for k, v in g.items()}


# ----------------------------------------------------------------------------- decoder loop and the whole model
# Verified on B200 against the reference's golden gradients (tests/test_train_step_gpu.py) and with fp64 stand-ins on the CPU.
@torch.no_grad()
def head_conv_train(x: torch.Tensor, conv: torch.nn.Conv3d):
    """A 3x3x3 convolution with fp32 logits and few output channels (mask heads :1380, final_block :1392)."""
    cw = _conv_pack(conv)
    logits = ops.conv3d(x, cw.w, cw.b, cw.cout, cw.k, pad=1, out_f32=True, w_tc=cw.w_tc)[0]
    return logits, dict(x=x, conv=conv)


@torch.no_grad()
def head_conv_backward(dlogits: torch.Tensor, saved: dict):
    """dlogits fp32 [B,h,w,d,Cout] -> (dx bf16, dW [Cout,Cin,3,3,3] fp32, dbias fp32).  The gradient is rounded to bf16 and
    zero-padded to a multiple of 8 channels (16-byte vectors of ltu_conv3d_wgrad); the padded filter rows are zero."""
    conv = saved["conv"]
    cout, cin, k = conv.weight.shape[0], conv.weight.shape[1], conv.weight.shape[2]
    cp = (cout + 7) // 8 * 8
    dy = torch.zeros(*dlogits.shape[:-1], cp, dtype=_ACT, device=dlogits.device)
    dy[..., :cout] = dlogits.to(_ACT)

    def build_padded():
        w = conv.weight.detach().new_zeros(cp, cin, k, k, k)
        w[:cout].copy_(conv.weight.detach())
        return SimpleNamespace(weight=w, bias=None, stride=(1, 1, 1))
    padded = _memo(("head_pad",), [conv.weight], build_padded)
    dx, dw, db = conv3d_backward(saved["x"], dy, padded)
    return dx, dw[:cout].contiguous(), db[:cout].contiguous()


@torch.no_grad()
def decoder_train(bottle: torch.Tensor, skips, dec, dim_output: int, drop: DropoutState = _NO_DROP):
    """ROIDecoder.forward (model/Unet_3Dblock.py:1359-1396) in training mode: returns (probs fp32 [B,C,H,W,D], mask_list,
    saved).  `dec` = lintransunet_b200.unet.ROIDecoder; bottle / skips from encoder_train."""
    from .unet import ROIBridge
    n = len(dec.num_layers)
    tr = dec.bridge_list[n - 1].transformer
    x, sv_bottle = transformer_stack_train(bottle, tr.layers, tr.pos_encoders[0], drop)
    levels, mask_list = [], []
    for i in range(1, n):
        fd = 2 if (n - i) % 2 == 0 else 1                                     # :1375-1378
        xu = ops.upsample_trilinear(x, fd)
        lvl = n - 1 - i
        logits, sv_m = head_conv_train(xu, dec.mask_conv_list[lvl])            # :1380
        mask, fg = ops.mask_softmax(logits, want_mask=True)
        mask_list.append(mask)
        skip, sv_g = gate_train(skips[-i], xu, dec.att_conv_list[lvl])         # :1384-1385
        bridge, sv_b = dec.bridge_list[lvl], None
        if isinstance(bridge, ROIBridge):
            skip, sv_b = roi_bridge_train(skip, fg, bridge, drop=drop)         # :1387-1388
        x, sv_u = upblock_train(xu, skip, dec.block_list[i - 1], drop)
        levels.append(dict(fd=fd, lvl=lvl, logits=logits, m=sv_m, g=sv_g, b=sv_b, u=sv_u))
    logits, sv_f = head_conv_train(x, dec.final_block)                         # :1392
    probs = ops.head_d2s_softmax(logits, dim_output, want_probs=True, want_onehot=False, want_labels=False)[0]
    return probs, mask_list, dict(bottle=sv_bottle, levels=levels, final=sv_f, final_logits=logits, n=n, cout=dim_output)


@torch.no_grad()
def decoder_backward(dprobs: torch.Tensor, dmask_list, saved: dict):
    """Returns (d_bottle, d_skips (one per encoder block), gradients keyed like ROIDecoder's state_dict)."""
    n = saved["n"]
    grads: Dict[str, torch.Tensor] = {}
    dlogits = ops.head_d2s_softmax_bwd(saved["final_logits"], dprobs.contiguous().float(), saved["cout"])
    dx, grads["final_block.weight"], grads["final_block.bias"] = head_conv_backward(dlogits, saved["final"])
    d_skips = [None] * (n - 1)
    for i in range(n - 1, 0, -1):
        lv = saved["levels"][i - 1]
        lvl = lv["lvl"]
        dxu, dskip, g = upblock_backward(dx, lv["u"])
        for k, v in g.items():
            grads[f"block_list.{i - 1}.{k}"] = v
        if lv["b"] is not None:
            dskip, g = roi_bridge_backward(dskip, lv["b"])
            for k, v in g.items():
                grads[f"bridge_list.{lvl}.{k}"] = v
        dskip, dup, g = gate_backward(dskip, lv["g"])
        for k, v in g.items():
            grads[f"att_conv_list.{lvl}.{k}"] = v
        dxu = dxu + dup
        if dmask_list is not None and dmask_list[i - 1] is not None:            # deep supervision on this mask head
            dlm = ops.mask_softmax_bwd(lv["logits"], dmask_list[i - 1].contiguous().float())
            dxm, grads[f"mask_conv_list.{lvl}.weight"], grads[f"mask_conv_list.{lvl}.bias"] = head_conv_backward(dlm, lv["m"])
            dxu = dxu + dxm
        d_skips[n - 1 - i] = dskip                                               # skips[-i]
        dx = ops.upsample_trilinear_bwd(dxu, lv["fd"])
    d_bottle, g = transformer_stack_backward(dx, saved["bottle"])
    for k, v in g.items():
        grads[f"bridge_list.{n - 1}.transformer." + k.replace("pos.", "pos_encoders.0.")] = v
    return d_bottle, d_skips, grads


def model_loss_and_gradients(model, x: torch.Tensor, masks: torch.Tensor, drop: DropoutState = _NO_DROP):
    """One training step's loss and parameter gradients of a binary MaskTransUnet on the bf16 path (dropout-free unless a
    DropoutState is passed):
    encoder_train -> decoder_train -> lintransunet_b200.losses.deep_supervision_loss (torch autograd on the outputs only)
    -> decoder_backward -> encoder_backward.  Returns (total, terms, grads keyed like model.state_dict())."""
    from . import losses
    with torch.no_grad():
        bottle, skips, sv_e = encoder_train(x, model.encode, drop)
        probs, mask_list, sv_d = decoder_train(bottle, skips, model.decode, model.dim_output, drop)
        drop.finish()
    p = probs.detach().requires_grad_(True)
    ms = [m.detach().requires_grad_(True) for m in mask_list]
    with torch.enable_grad():
        total, terms = losses.deep_supervision_loss(p, ms, masks)
        dall = torch.autograd.grad(total, [p] + ms)
    with torch.no_grad():
        d_bottle, d_skips, g_dec = decoder_backward(dall[0], list(dall[1:]), sv_d)
        g_enc = encoder_backward(d_bottle, d_skips, sv_e)
    grads = {f"encode.{k}": v for k, v in g_enc.items()}
    grads.update({f"decode.{k}": v for k, v in g_dec.items()})
    return total.detach(), terms, grads
