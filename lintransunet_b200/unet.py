"""Drop-in ``MaskTransUnet`` (reference model/trans_3DUnet.py:150-204) for B200.

The module tree below exists to reproduce the reference's constructor signature, submodule
names and ``state_dict`` layout (614 tensors for the default config): parameters live
in ordinary ``nn.Conv3d`` / ``nn.Linear`` / ``nn.LayerNorm`` containers whose own ``forward`` is
never called.  ``MaskTransUnet.forward`` runs the whole network through the sm_100a kernels of
``libltu_b200.so`` in a channels-last ``[B,H,W,D,C]`` layout; the only PyTorch compute calls are
the plain cuBLAS GEMMs of the ``nn.Linear`` layers.  There is no CPU path.

Precision: inside ``torch.autocast`` (any dtype; the reference scripts use
``torch.cuda.amp.autocast()``) activations are stored in bf16 with fp32 arithmetic and the
3x3x3 convolutions run on tcgen05 tensor cores; otherwise everything is fp32.  ``precision``
can force either.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _native, ops

__all__ = ["MaskTransUnet", "Encoder", "ROIDecoder", "Model_Dict", "get_model_dict"]


# ----------------------------------------------------------------------------- containers
class _Holder(nn.Module):
    """Parameter container: mirrors a reference submodule by name, never executed."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container: the computation runs in MaskTransUnet.forward")


class DownBlock(_Holder):
    """model/Unet_3Dblock.py:290-341."""

    def __init__(self, cin: int, cout: int, k: int, stride):
        super().__init__()
        self.conv1 = nn.Conv3d(cin, cin, k, stride=1, padding=k // 2)
        self.conv2 = nn.Conv3d(cin, cout, k, stride=stride, padding=k // 2)
        self.stride = tuple(stride)


class UpBlock(_Holder):
    """model/Unet_3Dblock.py:504-557."""

    def __init__(self, cin: int, cout: int, k: int):
        super().__init__()
        self.conv1 = nn.Conv3d(cin, cout, k, stride=1, padding=k // 2)
        self.conv2 = nn.Conv3d(2 * cout, cout, k, stride=1, padding=k // 2)


class Conv3dPosEmbedding(_Holder):
    """model/trans_block.py:70-96."""

    def __init__(self, dim: int):
        super().__init__()
        self.proj = nn.Conv3d(dim, dim, 3, stride=1, padding=1, groups=dim)


class MultihAttention(_Holder):
    """model/trans_block.py:127-166."""

    def __init__(self, d_model: int, nhead: int):
        super().__init__()
        assert d_model % nhead == 0
        self.nhead = nhead
        self.linears = nn.ModuleList([nn.Linear(d_model, d_model) for _ in range(4)])


class SelfAttentionLayer(_Holder):
    """model/trans_block.py:169-211."""

    def __init__(self, d_model: int, nhead: int):
        super().__init__()
        self.self_attn = MultihAttention(d_model, nhead)
        self.linear1 = nn.Linear(d_model, 2 * d_model)
        self.linear2 = nn.Linear(2 * d_model, d_model)
        self.layer_norm1 = nn.LayerNorm(d_model, eps=1e-6)
        self.layer_norm2 = nn.LayerNorm(d_model, eps=1e-6)


class _EmbedConv(_Holder):
    """DownEmbedBlock / UpEmbedBlock (model/Unet_3Dblock.py:343-432): module_list.0.<idx> is the conv."""

    def __init__(self, cin: int, cout: int, conv_index: int, stride: int):
        super().__init__()
        inner = [nn.Identity() for _ in range(conv_index)]
        inner.append(nn.Conv3d(cin, cout, 3, stride=stride, padding=1))
        self.module_list = nn.Sequential(nn.Sequential(*inner))

    @property
    def conv(self) -> nn.Conv3d:
        return self.module_list[0][-1]


class EmbedAttention3DBlock(_Holder):
    """model/Unet_3Dblock.py:435-501."""

    def __init__(self, in_dim: int, d_model: int, nhead: int, N: int):
        super().__init__()
        self.in_dim, self.d_model, self.nhead, self.N = in_dim, d_model, nhead, N
        self.down_embed = _EmbedConv(in_dim, d_model, conv_index=0, stride=2)
        self.up_embed = _EmbedConv(d_model, in_dim, conv_index=1, stride=1)
        self.pos_encoder = Conv3dPosEmbedding(d_model)
        self.layers = nn.ModuleList([SelfAttentionLayer(d_model, nhead) for _ in range(N)])


class PosAttention3DBlock(_Holder):
    """model/Unet_3Dblock.py:224-274 (only pos_encoders[0] is used, :267-270; 1..N-1 are dead
    parameters that stay in the state_dict)."""

    def __init__(self, d_model: int, nhead: int, N: int):
        super().__init__()
        self.d_model, self.nhead, self.N = d_model, nhead, N
        self.pos_encoders = nn.ModuleList([Conv3dPosEmbedding(d_model) for _ in range(N)])
        self.layers = nn.ModuleList([SelfAttentionLayer(d_model, nhead) for _ in range(N)])


class ConnectBridge(_Holder):
    """model/Unet_3Dblock.py:647-670."""

    def __init__(self, d_model: int, nhead: int, N: int):
        super().__init__()
        self.transformer = PosAttention3DBlock(d_model, nhead, N)


class ROIBridge(_Holder):
    """model/Unet_3Dblock.py:673-755; ROI constants :697-714."""

    def __init__(self, in_dim: int, d_model: int, nhead: int, N: int, roi_size: int, mask_threshold: float = 0.5):
        super().__init__()
        self.transformer = EmbedAttention3DBlock(in_dim, d_model, nhead, N)
        self.roi_size = roi_size
        self.h_roi_size = roi_size
        self.w_roi_size = int(roi_size * 0.6)
        self.eval_roi_size = int(1.2 * roi_size)
        self.eval_h_roi_size = self.eval_roi_size
        self.eval_w_roi_size = int(self.eval_h_roi_size * 0.6)
        self.min_h_roi = self.eval_roi_size // 2
        self.min_w_roi = self.eval_w_roi_size // 2
        self.mask_threshold = mask_threshold


class InitialBridge(_Holder):
    """model/Unet_3Dblock.py:1180-1199: identity, no parameters."""


class SpatialAttention3DBlock(_Holder):
    """model/Unet_3Dblock.py:194-221: three 1x1x1 convs."""

    def __init__(self, c_skip: int, c_up: int, c_inter: int):
        super().__init__()
        self.W_x = nn.Sequential(nn.Conv3d(c_skip, c_inter, 1))
        self.W_g = nn.Sequential(nn.Conv3d(c_up, c_inter, 1))
        self.psi = nn.Sequential(nn.Conv3d(c_inter, 1, 1))


class Encoder(_Holder):
    """model/Unet_3Dblock.py:560-607."""

    def __init__(self, num_layers: Sequence[int], dim_input: int, kernel_size: int = 3, dropout=None):
        super().__init__()
        self.num_layers = list(num_layers)
        self.block_list = nn.ModuleList([
            DownBlock(num_layers[i - 1], num_layers[i], kernel_size, (2, 2, (i - 1) % 2 + 1))
            for i in range(1, len(num_layers))])
        self.input_block = nn.Conv3d(dim_input * 4, num_layers[0], kernel_size, stride=1, padding=kernel_size // 2)


class ROIDecoder(_Holder):
    """model/Unet_3Dblock.py:1277-1396."""

    def __init__(self, num_layers: Sequence[int], roi_size_list: Sequence[int], is_roi_list: Sequence[bool],
                 dim_output: int, kernel_size: int = 3, nhead_lens: int = 32, dropout: float = 0.2, N: int = 8):
        super().__init__()
        L = list(num_layers)
        self.num_layers = L
        bridges: List[nn.Module] = []
        for i in range(len(L) - 1):
            if is_roi_list[i]:
                dm = min(4 * L[i], 256)
                bridges.append(ROIBridge(L[i], dm, dm // 32, N, roi_size_list[i]))
            else:
                bridges.append(InitialBridge())
        bridges.append(ConnectBridge(L[-1], L[-1] // nhead_lens, N))
        self.bridge_list = nn.ModuleList(bridges)
        self.mask_conv_list = nn.ModuleList([nn.Conv3d(L[i], dim_output, kernel_size, padding=kernel_size // 2)
                                             for i in range(1, len(L))])
        self.att_conv_list = nn.ModuleList([SpatialAttention3DBlock(L[i - 1], L[i], L[i - 1])
                                            for i in range(1, len(L))])
        self.block_list = nn.ModuleList([UpBlock(L[-i], L[-i - 1], kernel_size) for i in range(1, len(L))])
        self.final_block = nn.Conv3d(L[0], dim_output * 4, kernel_size, stride=1, padding=kernel_size // 2)


# ----------------------------------------------------------------------------- packed weights
class _ConvW:
    """Derived cache of one conv: fp32 [taps][Cin][Cout] for the CUDA-core kernel, bf16
    [Cout16][Kpad] (K-major, K = tap*Cin + c) for the tcgen05 kernel.  `cin_pad` appends zero input
    channels (the stem's 4 channels are padded to 8 = one 16-byte bf16 vector)."""

    def __init__(self, conv: nn.Conv3d, want_tc: bool, cin_pad: int = 0, fold_up2: bool = False,
                 aux: Optional[nn.Conv3d] = None, n_inputs: int = 1, f32_out: bool = False):
        w = conv.weight.detach()
        bias = conv.bias.detach() if conv.bias is not None else None
        self.n_aux = 0
        if aux is not None:          # fused head: extra output rows after the main ones (same input, same geometry)
            self.n_aux = aux.weight.shape[0]
            w = torch.cat([w, aux.weight.detach()], 0)
            bias = torch.cat([bias, aux.bias.detach()], 0)
        if cin_pad and cin_pad > w.shape[1]:
            w = torch.cat([w, w.new_zeros(w.shape[0], cin_pad - w.shape[1], *w.shape[2:])], 1)
        cout, cin, k = w.shape[0], w.shape[1], w.shape[2]
        self.k, self.cin, self.cout = k, cin, cout - self.n_aux        # cout = MAIN output channels
        self.w = w.permute(2, 3, 4, 1, 0).reshape(k * k * k, cin, cout).contiguous().float()
        self.b = bias.float().contiguous() if bias is not None else None
        self.w_tc = None
        if want_tc and (cin & (cin - 1)) == 0 and cin >= 8 and cout <= 256:
            ktot = k * k * k * cin
            kpad = (ktot + 63) // 64 * 64
            # rows padded to 16 (32 for the 3x3x3 layers the TMA-halo kernel can take; the other kernels read Cout16 rows)
            rpad = 32 if (k == 3 and cin >= 16) else 16
            wt = torch.zeros((cout + rpad - 1) // rpad * rpad, kpad, dtype=torch.bfloat16, device=w.device)
            wt[:cout, :ktot] = w.permute(0, 2, 3, 4, 1).reshape(cout, ktot).to(torch.bfloat16)
            self.w_tc = wt
        self.w_tc_fold = None
        if want_tc and fold_up2 and self.w_tc is not None and k == 3:
            self.w_tc_fold = _fold_up2_weights(w.float(), cout, cin)
        # small-channel stride-1 3x3x3 layers: super-voxel repacking for the TMA-halo tcgen05 kernel (ops.SVPack)
        self.sv = None
        ci = cin // n_inputs
        if want_tc and k == 3 and tuple(conv.stride) == (1, 1, 1) and ci in (8, 16, 32) and cin == ci * n_inputs and n_inputs in (1, 2):
            n_main = 0 if f32_out else self.cout
            n_aux = cout if f32_out else self.n_aux
            g = 64 // ci
            if g * n_main <= 128 and g * (n_main + n_aux) <= 256 and g * n_aux <= 64:
                self.sv = ops.sv_pack(w.float(), self.b, n_main, n_aux, ci, n_inputs)


def _fold_up2_weights(w: torch.Tensor, cout: int, cin: int) -> torch.Tensor:
    """nn.Upsample(nearest, x2) followed by a 3x3x3 conv == 8 parity-class 2x2x2 convs on the
    low-resolution input.  Per axis, output parity 0 reads source offsets (-1, 0) with weights
    (w0, w1+w2), parity 1 reads (0, +1) with (w0+w1, w2).  Returns bf16 [8][Cout16][Kpad],
    K index = (th*4+tw*2+td)*Cin + c (see ltu_conv3d_tc)."""
    f = w.new_zeros(2, 2, 3)                      # [parity][folded tap][original tap]
    f[0, 0, 0] = 1; f[0, 1, 1] = 1; f[0, 1, 2] = 1
    f[1, 0, 0] = 1; f[1, 0, 1] = 1; f[1, 1, 2] = 1
    # w: [Cout, Cin, kh, kw, kd] -> wf[pa,pb,pc, Cout, th,tw,td, Cin]
    wf = torch.einsum("oixyz,atx,buy,gvz->abgotuvi", w, f, f, f)
    ktot = 8 * cin
    kpad = (ktot + 63) // 64 * 64
    out = torch.zeros(8, (cout + 15) // 16 * 16, kpad, dtype=torch.bfloat16, device=w.device)
    out[:, :cout, :ktot] = wf.reshape(8, cout, ktot).to(torch.bfloat16)
    return out


class _LayerW:
    def __init__(self, layer: SelfAttentionLayer, dtype: torch.dtype):
        lin = layer.self_attn.linears
        c = lambda t: t.detach().to(dtype).contiguous()
        self.w_qkv = c(torch.cat([lin[0].weight, lin[1].weight, lin[2].weight], 0))
        self.b_qkv = c(torch.cat([lin[0].bias, lin[1].bias, lin[2].bias], 0))
        self.w_o, self.b_o = c(lin[3].weight), c(lin[3].bias)
        self.w_1, self.b_1 = c(layer.linear1.weight), c(layer.linear1.bias)
        self.w_2, self.b_2 = c(layer.linear2.weight), c(layer.linear2.bias)
        f = lambda t: t.detach().float().contiguous()
        self.fused = dtype == torch.bfloat16
        if self.fused:                                 # operands of the fused tcgen05 Linear layers
            self.wt_o, self.bf_o = ops.pack_linear_tc(lin[3].weight), f(lin[3].bias)
            self.wt_1, self.bf_1 = ops.pack_linear_tc(layer.linear1.weight), f(layer.linear1.bias)
            self.wt_2, self.bf_2 = ops.pack_linear_tc(layer.linear2.weight), f(layer.linear2.bias)
        # fused FFN kernel (ltu_ffn_fused): bf16 weights in the nn.Linear layout, fp32 biases
        self.ffn = self.fused and ops.ffn_fused_supported(layer.linear1.in_features)
        if self.ffn:
            self.b1_f32, self.b2_f32 = f(layer.linear1.bias), f(layer.linear2.bias)
        # fused query half (ltu_attn_out_fused): K/V projection stays a cuBLAS GEMM feeding kv_reduce
        self.attn_fused = self.fused and ops.attn_out_fused_supported(layer.linear1.in_features, layer.self_attn.nhead)
        if self.attn_fused:
            self.w_kv = c(torch.cat([lin[1].weight, lin[2].weight], 0))
            self.b_kv = c(torch.cat([lin[1].bias, lin[2].bias], 0))
            self.w_q, self.bq_f32, self.bo_f32 = c(lin[0].weight), f(lin[0].bias), f(lin[3].bias)
        self.g1, self.be1 = f(layer.layer_norm1.weight), f(layer.layer_norm1.bias)
        self.g2, self.be2 = f(layer.layer_norm2.weight), f(layer.layer_norm2.bias)
        self.nhead = layer.self_attn.nhead
        # TMA + tcgen05 Linear layers with fused epilogues (ltu_linear_fused): d_model 256 -> every Linear of the layer;
        # d_model 128 -> the K/V projection.  The weights are the nn.Linear matrices themselves (bf16), biases fp32.
        d = layer.linear1.in_features
        self.lin = self.fused and d == 256 and ops.linear_fused_supported(d, 3 * d)
        if self.lin:
            self.bqkv_f32, self.bo_f32 = f(self.b_qkv), f(lin[3].bias)
            self.b1_f32, self.b2_f32 = f(layer.linear1.bias), f(layer.linear2.bias)
        self.lin_kv = self.attn_fused and ops.linear_fused_supported(d, 2 * d)
        if self.lin_kv:
            self.bkv_f32 = f(self.b_kv)
        # fp32 path: W^T as [1][K][N] for the CUDA-core GEMM (ops.linear_fma)
        self.fma = dtype == torch.float32
        if self.fma:
            kn = lambda w: w.detach().float().t().contiguous().unsqueeze(0)
            self.wp_qkv, self.wp_o, self.wp_1, self.wp_2 = kn(self.w_qkv), kn(self.w_o), kn(self.w_1), kn(self.w_2)


def _pos_w(pe: Conv3dPosEmbedding):
    # reference applies the depthwise conv on the (D,H,W)-permuted view: kernel axes = (kd,kh,kw);
    # native [kh,kw,kd][C] packing
    w = pe.proj.weight.detach()[:, 0]                      # [C, kd, kh, kw]
    return (w.permute(2, 3, 1, 0).reshape(27, -1).contiguous().float(),
            pe.proj.bias.detach().float().contiguous())


class _Plan:
    """All derived weights for one (device, precision).  Rebuilt when any parameter changes
    (load_state_dict / .to() / optimizer step bump the version or the storage)."""

    def __init__(self, model: "MaskTransUnet", dtype: torch.dtype):
        tc = dtype == torch.bfloat16
        enc, dec = model.encode, model.decode
        self.dtype = dtype
        self.stem = _ConvW(enc.input_block, tc, cin_pad=8 if tc else 0)
        self.down = [(_ConvW(b.conv1, tc), _ConvW(b.conv2, tc), b.stride) for b in enc.block_list]
        self.mask = [_ConvW(c, tc, f32_out=True) for c in dec.mask_conv_list]
        self.gate = []
        for a in dec.att_conv_list:
            psi = a.psi[0]
            self.gate.append((_ConvW(a.W_x[0], tc), _ConvW(a.W_g[0], tc),
                              psi.weight.detach().reshape(-1).float().contiguous(),
                              psi.bias.detach().float().contiguous()))
        self.up = [(_ConvW(b.conv1, tc), _ConvW(b.conv2, tc, n_inputs=2)) for b in dec.block_list]
        # bf16 path: the mask head of level i reads the same upsampled x as UpBlock.conv1 -> one fused launch
        n = len(dec.block_list)
        self.up_dual = [_ConvW(dec.block_list[i].conv1, True, aux=dec.mask_conv_list[n - 1 - i]) for i in range(n)] if tc else None
        self.final = _ConvW(dec.final_block, tc, f32_out=True)
        self.bridges: List[Optional[dict]] = []
        for br in dec.bridge_list:
            if isinstance(br, ROIBridge):
                t = br.transformer
                self.bridges.append(dict(kind="roi", down=_ConvW(t.down_embed.conv, tc), up=_ConvW(t.up_embed.conv, tc, fold_up2=True),
                                         pos=_pos_w(t.pos_encoder), layers=[_LayerW(l, dtype) for l in t.layers],
                                         mod=br))
            elif isinstance(br, ConnectBridge):
                t = br.transformer
                self.bridges.append(dict(kind="bottle", pos=_pos_w(t.pos_encoders[0]),
                                         layers=[_LayerW(l, dtype) for l in t.layers]))
            else:
                self.bridges.append(None)


def _named_weights(model: nn.Module):
    """(name, tensor) of every weight the forward reads, in `named_parameters()` order.  On an nn.DataParallel replica
    `_parameters` is empty: torch's replicate() re-attaches the broadcast copies as plain attributes and lists them in
    `_former_parameters`, which is what the derived-weight cache and the training path have to look at."""
    out = []
    for prefix, mod in model.named_modules():
        params = mod._parameters
        if not params:
            params = getattr(mod, "_former_parameters", None) or {}
        for k, p in params.items():
            if p is not None:
                out.append((f"{prefix}.{k}" if prefix else k, p))
    return out


def _signature(model: nn.Module):
    return tuple((p.data_ptr(), p._version) for _, p in _named_weights(model))


# ----------------------------------------------------------------------------- training
class _NativeTrainFunction(torch.autograd.Function):
    """``(probs, *mask_list) = f(x, *parameters)`` with the native forward (training mode) and the native backward of
    lintransunet_b200/backward.py, so that ``loss.backward()`` of the reference's train step
    (utils/utils_3D_embed_full.py:63-91) fills ``p.grad`` of every parameter.  bf16 activations, fp32 gradients.
    The functions it calls reproduce the reference's gradients on B200 (tests/test_train_step_gpu.py); the wrapper is
    checked against them on the GPU (tests/test_train_step_gpu.py::test_autograd_wiring_fills_param_grads) and with fp64
    stand-in kernels on the CPU (tests/test_backward_composition_cpu.py).  Training-mode dropout: one
    backward.DropoutState per forward, masks re-generated in the backward."""

    @staticmethod
    def forward(ctx, model, x, *params):
        from . import backward as BW
        drop = BW.DropoutState(model.dropout, x.device)
        bottle, skips, sv_e = BW.encoder_train(x, model.encode, drop)
        probs, mask_list, sv_d = BW.decoder_train(bottle, skips, model.decode, model.dim_output, drop)
        drop.finish()
        ctx.saved_state = (sv_e, sv_d)
        ctx.names = [n for n, _ in _named_weights(model)]
        ctx.dtypes = [p.dtype for p in params]
        ctx.out_shapes = [tuple(probs.shape)] + [tuple(m.shape) for m in mask_list]
        return (probs, *mask_list)

    @staticmethod
    def backward(ctx, dprobs, *dmasks):
        from . import backward as BW
        sv_e, sv_d = ctx.saved_state
        if dprobs is None:
            dprobs = torch.zeros(ctx.out_shapes[0], dtype=torch.float32, device=sv_d["final_logits"].device)
        d_bottle, d_skips, g_dec = BW.decoder_backward(dprobs, list(dmasks), sv_d)
        g_enc = BW.encoder_backward(d_bottle, d_skips, sv_e)
        grads = {f"encode.{k}": v for k, v in g_enc.items()}
        grads.update({f"decode.{k}": v for k, v in g_dec.items()})
        out = []
        for name, dt in zip(ctx.names, ctx.dtypes):
            g = grads.get(name)                      # None: dead parameters (pos_encoders.1-7) and unsupervised heads
            out.append(None if g is None else g.to(dt))
        return (None, None, *out)


# ----------------------------------------------------------------------------- the model
class MaskTransUnet(nn.Module):
    """Same constructor, forward contract and state_dict as the reference
    (model/trans_3DUnet.py:161-204): train mode returns ``(probs, mask_list)``, eval mode the
    one-hot argmax."""

    def __init__(self, num_layers: list, roi_size_list: list, is_roi_list: list, dim_input: int, dim_output: int,
                 kernel_size: int = 3, dropout: float = 0.3):
        super().__init__()
        if kernel_size != 3:
            raise ValueError("only kernel_size=3 (the reference default) is supported")
        if dim_input != 1:
            raise ValueError("windows_embedding requires dim_input == 1 (model/Unet_3Dblock.py:132)")
        self.num_layers = list(num_layers)
        self.kernel_size = kernel_size
        self.dropout = dropout
        self.dim_input = dim_input
        self.dim_output = dim_output
        self.roi_size_list = list(roi_size_list)
        self.is_roi_list = list(is_roi_list)
        self.encode = Encoder(self.num_layers, dim_input, kernel_size, dropout)
        self.decode = ROIDecoder(self.num_layers, self.roi_size_list, self.is_roi_list, dim_output,
                                 dropout=dropout)
        # runtime knobs (not part of the reference API)
        self.precision: Optional[str] = None          # None = follow autocast, or "fp32" / "bf16"
        self.use_tensor_cores = os.environ.get("LTU_DISABLE_TC", "0") != "1"
        self.forced_boxes: Optional[Dict[int, torch.Tensor]] = None   # tests: teacher-forced ROI boxes
        self.record: Optional[dict] = None            # tests: set to {} to capture taps (channels-last)
        # the forward has no host sync (the ROI boxes stay on the device), so an inference forward of
        # a fixed input shape is captured once into a CUDA graph and replayed (~600 launches -> 1)
        self.use_cuda_graphs = os.environ.get("LTU_CUDA_GRAPHS", "1") != "0"
        # bf16 path, EXPERIMENTAL (off): nn.Linear + bias + GELU / residual + LayerNorm in one tcgen05 launch each.
        # Correct (tests/test_ops_gpu.py::test_linear_tc_fused_epilogues) but 4.4x slower than cuBLAS + the separate
        # bandwidth-bound kernels at the model's shapes (profiles/r1_conv_variants.md) until the GEMM keeps its
        # weights resident and widens the epilogue.
        self.use_fused_linear = os.environ.get("LTU_FUSED_LINEAR", "0") == "1"
        # Fused FFN half of the encoder layers with d_model = 128 (bridge 1: 57 408 tokens per sample at 128^3):
        # 155 us vs 306 us for cuBLAS + gelu + cuBLAS + add_layernorm at batch 8.  LTU_FUSED_FFN=0 is the A/B switch.
        self.use_fused_ffn = os.environ.get("LTU_FUSED_FFN", "1") == "1"
        # Fused query half of the same layers (Q projection + readout + output projection + residual + LayerNorm1):
        # 253 us vs 318 us for the attention half at batch 8.  LTU_FUSED_ATTN=0 is the A/B switch.
        self.use_fused_attn = os.environ.get("LTU_FUSED_ATTN", "1") == "1"
        # bf16 path: compute the mask head inside UpBlock.conv1's launch (same input), fp32 logits as a second output
        self.fuse_mask_head = os.environ.get("LTU_FUSE_MASK_HEAD", "1") != "0"
        # inference: do not launch a mask head nobody reads (level without a ROI bridge); the outputs are bit-identical
        # either way (tests/test_model_gpu.py::test_dead_mask_head_does_not_change_the_result).  LTU_KEEP_DEAD_HEAD=1 = A/B.
        self.skip_dead_mask_head = os.environ.get("LTU_KEEP_DEAD_HEAD", "0") != "1"
        # bf16 path: carry the residual stream of the encoder layers that run as separate kernels (d_model 256) as two
        # bf16 words (hi + lo): the bf16 rounding of the LayerNorm outputs is the largest single term of the bf16
        # path's error (4.6e-2 -> 3.5e-2 on the 64x64x16 case); the reference keeps it in fp32.
        self.split_token_stream = os.environ.get("LTU_SPLIT_STREAM", "1") == "1"
        # bf16 path: the nn.Linear layers of the d_model-256 encoder layers (and bridge 1's K/V projection) on the native
        # TMA + tcgen05 GEMM with fused epilogues instead of cuBLAS + gelu + add_layernorm.  LTU_NATIVE_LINEAR=0 = A/B.
        self.use_native_linear = os.environ.get("LTU_NATIVE_LINEAR", "1") == "1"
        # d_model 128: reduce K and V inside the K/V projection's epilogue (ltu_kv_project_reduce).  LTU_KV_PROJECT=0 = A/B.
        self.fuse_kv_project = os.environ.get("LTU_KV_PROJECT", "1") == "1"
        self.fold_readout = os.environ.get("LTU_FOLD_READOUT", "1") == "1"      # d_model 256: q_readout folded into the GEMMs around it
        # training: autograd through the native backward (bf16 path; loss.backward() fills p.grad).  LTU_NATIVE_BACKWARD=0
        # turns a training forward with grad into an error instead (there is no other backward).
        self.native_backward = os.environ.get("LTU_NATIVE_BACKWARD", "1") == "1"
        self.max_cached_graphs = 16                   # one graph (+ private memory pool) per input shape, head and slot, LRU
        self.graph_slot = 0                           # which instance of a shape's graph _run replays (sliding_window.py)
        self._plans: Dict[tuple, tuple] = {}
        self._graphs: Dict[tuple, dict] = {}

    # -- pickling (the reference saves whole modules: torch.save(model, ...), train3D.py:291) ------------------------------
    def __getstate__(self):
        state = self.__dict__.copy()
        state["_plans"], state["_graphs"] = {}, {}        # derived weights and CUDA graphs are rebuilt on first use
        state["record"], state["forced_boxes"] = None, None
        return state

    # -- derived-weight cache ------------------------------------------------------------
    def _plan(self, device: torch.device, dtype: torch.dtype) -> _Plan:
        key = (device.index, dtype)
        sig = _signature(self)
        hit = self._plans.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        plan = _Plan(self, dtype)
        self._plans[key] = (sig, plan)
        return plan

    def _compute_dtype(self) -> torch.dtype:
        if self.precision is not None:
            if self.precision not in ("fp32", "bf16"):
                raise ValueError("precision must be None, 'fp32' or 'bf16'")
            return torch.float32 if self.precision == "fp32" else torch.bfloat16
        return torch.bfloat16 if torch.is_autocast_enabled() else torch.float32

    # -- CUDA graph replay of the inference heads ----------------------------------------------
    def _run(self, x: torch.Tensor, head: str):
        """Run `_forward_impl` eagerly or through a captured CUDA graph (eval / labels / logits)."""
        dtype = self._compute_dtype()
        x = x.contiguous().float()
        with torch.autocast("cuda", enabled=False):
            plan = self._plan(x.device, dtype)
            graphable = (self.use_cuda_graphs and head != "train" and self.record is None
                         and self.forced_boxes is None and not getattr(self, "_is_replica", False)
                         and not torch.cuda.is_current_stream_capturing())
            if not graphable:
                return self._forward_impl(x, plan, head)
            # graph_slot: the sliding-window driver replays two instances of the same forward on two streams (two static
            # input / output buffers, two memory pools), so that the latency-bound layers of one batch run under the other's
            key = (x.device.index, dtype, head, tuple(x.shape), int(getattr(self, "graph_slot", 0))) + self._knobs()
            ent = self._graphs.get(key)
            if ent is None or ent["plan"] is not plan:
                ent = self._capture(x, plan, head)
                self._graphs.pop(key, None)
                while len(self._graphs) >= self.max_cached_graphs:      # least recently used shape goes first
                    self._graphs.pop(next(iter(self._graphs)))
            else:
                self._graphs.pop(key)
            self._graphs[key] = ent                                     # (re)insert as most recently used
            ent["x"].copy_(x)
            ent["graph"].replay()
            _native.note_replayed(ent["launches"])       # kernels executed by the replay (bench gpu_launches)
            return ent["out"]

    def _knobs(self) -> tuple:
        """Every runtime switch that changes the launched kernels (part of the CUDA-graph cache key)."""
        return (self.use_tensor_cores, self.use_fused_linear, self.fuse_mask_head, self.use_fused_ffn, self.use_fused_attn,
                self.split_token_stream, self.use_native_linear, self.fuse_kv_project, self.fold_readout, self.skip_dead_mask_head, ops.USE_HALO_CONV, ops.USE_TC3_CONV, ops.USE_SV_CONV, ops.USE_CONCAT_TC3)

    def _capture(self, x: torch.Tensor, plan: "_Plan", head: str) -> dict:
        static_x = x.clone()
        side = torch.cuda.Stream(device=x.device)
        side.wait_stream(torch.cuda.current_stream(x.device))
        with torch.cuda.stream(side):                 # warm-up outside the capture (lazy inits, cuBLAS workspaces)
            self._forward_impl(static_x, plan, head)
        torch.cuda.current_stream(x.device).wait_stream(side)
        torch.cuda.synchronize(x.device)
        graph = torch.cuda.CUDAGraph()
        n0 = _native.lib().ltu_launch_count()
        with torch.cuda.graph(graph):
            out = self._forward_impl(static_x, plan, head)
        captured = int(_native.lib().ltu_launch_count() - n0)      # native kernels recorded in the graph
        _native.note_replayed(-captured)                          # capture itself executed nothing
        return dict(graph=graph, x=static_x, out=out, plan=plan, launches=captured)

    # -- building blocks -----------------------------------------------------------------
    def _tap(self, name: str, t: torch.Tensor):
        if self.record is not None:
            self.record[name] = t

    def _conv(self, x, cw: _ConvW, **kw):
        tc = self.use_tensor_cores
        return ops.conv3d(x, cw.w, cw.b, cw.cout, cw.k, w_tc=cw.w_tc if tc else None,
                          w_tc_fold=cw.w_tc_fold if tc else None, sv=cw.sv if tc else None, **kw)

    def _conv_in_act(self, x, cw: _ConvW, stride=(1, 1, 1), residual=None, x1=None, up2=False):
        """Conv3d -> InstanceNorm3d -> LeakyReLU (+ residual)."""
        y, partials, _ = self._conv(x, cw, stride=stride, pad=cw.k // 2, x1=x1, up2=up2, want_stats=True)
        V = y.shape[1] * y.shape[2] * y.shape[3]
        stats = ops.instnorm_finalize(partials, V)
        return ops.instnorm_apply(y, stats, ops.ACT_LRELU, residual=residual, inplace=True)

    def _encoder_layer(self, t, lw: _LayerW, lo=None, split=False):
        """SelfAttentionLayer.forward (model/trans_block.py:203-211) on tokens [B,N,C].  Returns (t, lo): with `split`
        (bf16 path, layers that run as separate kernels) the residual stream is carried as t + lo, two bf16 tensors
        (ops.add_layernorm_split): the Linear layers read t, the residual adds see 16 significant bits like the
        reference's fp32 LayerNorm outputs under autocast."""
        B, N, C = t.shape
        if lw.attn_fused and self.use_fused_attn and not (lw.fused and self.use_fused_linear):
            # K/V projection (cuBLAS) -> kv_reduce -> ONE kernel for Q projection, readout, output projection,
            # residual and LayerNorm1; then ONE kernel for the feed-forward half
            # fold_readout: the merge kernel of the context reduction also writes W_b = blockdiag(ctx_b) Wo^T and the fused
            # kernel runs two chained GEMMs per tile (Q projection, P W_b^T) instead of three
            w_o = lw.w_o if self.fold_readout else None
            if (lw.lin_kv and self.use_native_linear and self.fuse_kv_project and B <= 30
                    and ops.kv_project_reduce_supported(C, lw.nhead, N)):
                # K/V projection and context reduction in ONE launch: K and V never reach memory
                ctx = ops.kv_project_reduce(t, lw.w_kv, lw.bkv_f32, lw.nhead, w_o=w_o)
            else:
                if lw.lin_kv and self.use_native_linear:
                    kv = ops.linear_fused(t, lw.w_kv, lw.bkv_f32)
                else:
                    kv = F.linear(t, lw.w_kv, lw.b_kv)
                ctx = ops.kv_reduce(kv[..., :C], kv[..., C:], lw.nhead, w_o=w_o)
            if self.fold_readout:
                t = ops.attn_out_fused_w(t, lw.w_q, lw.bq_f32, ctx[1], lw.bo_f32, lw.g1, lw.be1, lw.nhead)
            else:
                t = ops.attn_out_fused(t, lw.w_q, lw.bq_f32, ops.ctx_pack(ctx), lw.w_o, lw.bo_f32, lw.g1, lw.be1, lw.nhead)
            if lw.ffn and self.use_fused_ffn:
                return ops.ffn_fused(t, lw.w_1, lw.b1_f32, lw.w_2, lw.b2_f32, lw.g2, lw.be2, 1e-6), None
            f = ops.gelu_(F.linear(t, lw.w_1, lw.b_1))
            f = F.linear(f, lw.w_2, lw.b_2)
            return ops.add_layernorm(t, f, lw.g2, lw.be2, 1e-6), None
        if lw.lin and self.use_native_linear:
            # d_model 256: four TMA + tcgen05 launches carry every Linear of the layer with its bias, GELU and
            # residual + LayerNorm (the split stream rides the operand ring; ltu_linear_fused)
            if self.fold_readout:
                # the query half of linear_attention rides the GEMMs around it: the QKV launch writes softmax(Q) / sqrt(32),
                # and the output projection runs with the sample's W_b = blockdiag(ctx_b) Wo^T -- (P ctx_b) Wo^T = P (ctx_b Wo^T):
                # q_readout never runs, the attention output is never written
                qkv = ops.linear_fused(t, lw.w_qkv, lw.bqkv_f32, softmax_cols=C)
                _, w_b = ops.kv_reduce(qkv[..., C:2 * C], qkv[..., 2 * C:], lw.nhead, w_o=lw.w_o)
                t, lo = ops.linear_fused(qkv, w_b, lw.bo_f32, ops.EPI_RES_LN, t, lo, lw.g1, lw.be1, 1e-6,
                                         want_lo=split, x_cols=C)             # P = columns [0, C) of the QKV rows
            else:
                qkv = ops.linear_fused(t, lw.w_qkv, lw.bqkv_f32)
                ctx = ops.kv_reduce(qkv[..., C:2 * C], qkv[..., 2 * C:], lw.nhead)
                att = ops.q_readout(qkv[..., :C], ctx, lw.nhead)
                t, lo = ops.linear_fused(att, lw.w_o, lw.bo_f32, ops.EPI_RES_LN, t, lo, lw.g1, lw.be1, 1e-6, want_lo=split)
            f = ops.linear_fused(t, lw.w_1, lw.b1_f32, ops.EPI_GELU)
            return ops.linear_fused(f, lw.w_2, lw.b2_f32, ops.EPI_RES_LN, t, lo, lw.g2, lw.be2, 1e-6, want_lo=split)
        if lw.fma and self.use_native_linear:
            # fp32 path: exact-FMA GEMMs on CUDA cores (no TF32, no library call)
            qkv = ops.linear_fma(t, lw.wp_qkv, lw.b_qkv, 3 * C)
            ctx = ops.kv_reduce(qkv[..., C:2 * C], qkv[..., 2 * C:], lw.nhead)
            att = ops.q_readout(qkv[..., :C], ctx, lw.nhead)
            t = ops.add_layernorm(t, ops.linear_fma(att, lw.wp_o, lw.b_o, C), lw.g1, lw.be1, 1e-6)
            f = ops.gelu_(ops.linear_fma(t, lw.wp_1, lw.b_1, 2 * C))
            return ops.add_layernorm(t, ops.linear_fma(f, lw.wp_2, lw.b_2, C), lw.g2, lw.be2, 1e-6), None
        qkv = F.linear(t, lw.w_qkv, lw.b_qkv)                           # LTU_NATIVE_LINEAR=0: cuBLAS, the A/B baseline
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
        ctx = ops.kv_reduce(k, v, lw.nhead)
        att = ops.q_readout(q, ctx, lw.nhead)
        if lw.fused and self.use_fused_linear:
            # O-projection + residual + LayerNorm1, FFN1 + GELU, FFN2 + residual + LayerNorm2: three launches
            t = ops.linear_tc(att, lw.wt_o, lw.bf_o, C, ops.EPI_RES_LN, residual=t, gamma=lw.g1, beta=lw.be1)
            f = ops.linear_tc(t, lw.wt_1, lw.bf_1, 2 * C, ops.EPI_GELU)
            return ops.linear_tc(f, lw.wt_2, lw.bf_2, C, ops.EPI_RES_LN, residual=t, gamma=lw.g2, beta=lw.be2), None
        o = F.linear(att, lw.w_o, lw.b_o)
        if split:
            t, lo = ops.add_layernorm_split(t, lo, o, lw.g1, lw.be1, 1e-6)
        else:
            t = ops.add_layernorm(t, o, lw.g1, lw.be1, 1e-6)
        if lw.ffn and self.use_fused_ffn:
            # linear1 + GELU + linear2 + residual + LayerNorm2 in one launch; the hidden activation stays on the SM
            return ops.ffn_fused(t, lw.w_1, lw.b1_f32, lw.w_2, lw.b2_f32, lw.g2, lw.be2, 1e-6), None
        f = ops.gelu_(F.linear(t, lw.w_1, lw.b_1))
        f = F.linear(f, lw.w_2, lw.b_2)
        if split:
            return ops.add_layernorm_split(t, lo, f, lw.g2, lw.be2, 1e-6)
        return ops.add_layernorm(t, f, lw.g2, lw.be2, 1e-6), None

    def _transformer(self, x, br: dict, name: str):
        """The 8-layer stack of PosAttention3DBlock / EmbedAttention3DBlock
        (model/Unet_3Dblock.py:265-270, :484-490) on a channels-last volume."""
        B, H, W, D, C = x.shape
        t = x.reshape(B, H * W * D, C)
        lo = None
        for i, lw in enumerate(br["layers"]):
            split = (self.split_token_stream and t.dtype == torch.bfloat16 and not (lw.fused and self.use_fused_linear)
                     and not (lw.attn_fused and self.use_fused_attn) and not (lw.ffn and self.use_fused_ffn))
            t, lo = self._encoder_layer(t, lw, lo, split)
            if i == 0:
                if split:
                    t, lo = ops.posenc_dwconv3_split(t.reshape(B, H, W, D, C), None if lo is None else lo.reshape(B, H, W, D, C),
                                                     br["pos"][0], br["pos"][1])
                    t, lo = t.reshape(B, H * W * D, C), lo.reshape(B, H * W * D, C)
                else:
                    t = ops.posenc_dwconv3(t.reshape(B, H, W, D, C), br["pos"][0], br["pos"][1]).reshape(B, H * W * D, C)
        return t.reshape(B, H, W, D, C)

    def _roi_bridge(self, skip, fg, br: dict, idx: int):
        """ROIBridge.forward (model/Unet_3Dblock.py:717-755)."""
        m: ROIBridge = br["mod"]
        B, h, w, d, C = skip.shape
        if self.forced_boxes is not None and idx in self.forced_boxes:
            box = self.forced_boxes[idx].to(device=skip.device, dtype=torch.float32).contiguous()
        else:
            box = ops.roi_bbox(fg, m.min_h_roi, m.min_w_roi, m.mask_threshold)
        self._tap(f"box{idx}", box)
        geo = (m.h_roi_size, m.w_roi_size, m.eval_h_roi_size, m.eval_w_roi_size)
        roi = ops.roi_resample(skip, box, (h, w), *geo, direction=0)
        self._tap(f"roi_in{idx}", roi)
        t = self._conv_in_act(roi, br["down"], stride=(2, 2, 2))
        t = self._transformer(t, br, f"bridge{idx}")
        t = self._conv_in_act(t, br["up"], up2=True)
        if t.shape[1:4] != roi.shape[1:4]:
            raise RuntimeError(f"ROI bridge {idx}: up_embed output {tuple(t.shape)} does not match the ROI "
                               f"{tuple(roi.shape)} (depth must be even at this level)")
        self._tap(f"roi_out{idx}", t)
        return ops.roi_resample(t, box, (h, w), *geo, direction=1)

    # -- forward -------------------------------------------------------------------------
    def forward(self, x: torch.Tensor):
        if not x.is_cuda:
            raise RuntimeError("lintransunet_b200.MaskTransUnet runs on CUDA (sm_100a) only; there is no CPU path")
        if self.training:
            # the reference's train-mode forward (trans_3DUnet.py:181-195) with its dropout (p = self.dropout) and, when a
            # parameter requires grad, autograd through the native backward
            want_grad = torch.is_grad_enabled() and any(p.requires_grad for _, p in _named_weights(self))
            if want_grad or (self.dropout and self.dropout > 0):
                if want_grad and not self.native_backward:
                    raise NotImplementedError("model.native_backward is off (LTU_NATIVE_BACKWARD=0): a training forward "
                                              "with grad needs it; wrap the call in torch.no_grad() or switch it on")
                if self._compute_dtype() != torch.bfloat16:
                    raise NotImplementedError("training (native backward, dropout) exists on the bf16 path only: call the "
                                              "model inside torch.autocast, as the reference's train step does "
                                              "(utils/utils_3D_embed_full.py:64), or set dropout=0.0 and use no_grad")
                with torch.autocast("cuda", enabled=False):
                    return self._forward_train(x, want_grad)
        B, _, H, W, D = x.shape
        if H % 32 or W % 32 or D % 4:
            raise ValueError("H and W must be multiples of 32 and D a multiple of 4")
        with torch.no_grad():
            out = self._run(x, "train" if self.training else "eval")
            if not self.training and self.use_cuda_graphs:
                out = out.clone()          # the graph's output buffer is reused by the next call
        return out

    def _forward_train(self, x: torch.Tensor, want_grad: bool = True):
        """Training-mode forward (dropout active) with autograd through the native backward: returns (probs, mask_list)
        like the reference's train-mode forward (model/trans_3DUnet.py:181-195)."""
        B, _, H, W, D = x.shape
        if H % 32 or W % 32 or D % 4:
            raise ValueError("H and W must be multiples of 32 and D a multiple of 4")
        xin = x.contiguous().float()
        if not want_grad:
            from . import backward as BW
            with torch.no_grad():
                drop = BW.DropoutState(self.dropout, x.device)
                bottle, skips, _ = BW.encoder_train(xin, self.encode, drop)
                probs, mask_list, _ = BW.decoder_train(bottle, skips, self.decode, self.dim_output, drop)
                drop.finish()
            return probs, mask_list
        out = _NativeTrainFunction.apply(self, xin, *[p for _, p in _named_weights(self)])
        return out[0], list(out[1:])

    def _forward_impl(self, x: torch.Tensor, P: _Plan, head: str):
        n = len(self.num_layers)
        # ---- Encoder.forward (model/Unet_3Dblock.py:596-607)
        a = ops.s2d_input(x, P.dtype, cpad=P.stem.cin)          # 8 (zero-padded) on the tensor-core path
        a = self._conv_in_act(a, P.stem)
        skips = []
        for i, (c1, c2, stride) in enumerate(P.down):
            s = self._conv_in_act(a, c1, residual=a)                 # DownBlock :327-331
            skips.append(s)
            self._tap(f"skip{i}", s)
            a = self._conv_in_act(s, c2, stride=stride)             # :335-336
        self._tap("bottle", a)
        # ---- ROIDecoder.forward (:1359-1396)
        x = self._transformer(a, P.bridges[n - 1], "bridge_bottle")
        self._tap(f"bridge{n-1}", x)
        mask_list = []
        for i in range(1, n):
            x = ops.upsample_trilinear(x, 2 if (n - i) % 2 == 0 else 1)          # :1375-1378
            lvl = n - 1 - i
            # The mask head of a level feeds (a) the deep-supervision loss, through mask_list, in training mode and (b) the
            # ROI box of that level's bridge (:1387).  A level without a ROI bridge (is_roi_list False: the finest one in
            # the reference configuration) has no consumer in an inference forward, whose result (:199-201) is the argmax
            # of the final block only: that head is dead code there and is not launched.
            need_mask = head == "train" or P.bridges[lvl] is not None or not self.skip_dead_mask_head
            fused = P.up_dual is not None and self.use_tensor_cores and self.fuse_mask_head and need_mask
            fg = None
            if fused:    # mask head (:1380) + UpBlock.conv1 (:547) share their input: one conv, two outputs
                cw = P.up_dual[i - 1]
                raw1, part1, _, logits = ops.conv3d(x, cw.w, cw.b, cw.cout, cw.k, pad=1, want_stats=True, w_tc=cw.w_tc,
                                                    n_aux=cw.n_aux, sv=cw.sv)
            elif need_mask:
                logits, _, _ = self._conv(x, P.mask[lvl], pad=1, out_f32=True)   # :1380
            if need_mask:
                mask, fg = ops.mask_softmax(logits, want_mask=(head == "train"))
                if mask is not None:
                    mask_list.append(mask)
            skip = skips[-i]
            wx, wg, psi_w, psi_b = P.gate[lvl]
            ga, pa, _ = self._conv(skip, wx, pad=0, want_stats=True)             # SpatialAttention3DBlock :217-221
            gg, pg, _ = self._conv(x, wg, pad=0, want_stats=True)
            V = skip.shape[1] * skip.shape[2] * skip.shape[3]
            skip = ops.gate_fused(ga, ops.instnorm_finalize(pa, V), gg, ops.instnorm_finalize(pg, V),
                                  psi_w, psi_b, skip)                            # :1384-1385
            br = P.bridges[lvl]
            if br is not None:
                skip = self._roi_bridge(skip, fg, br, lvl)                       # :1387-1388
            self._tap(f"bridge{lvl}", skip)
            c1, c2 = P.up[i - 1]
            if fused:
                V1 = raw1.shape[1] * raw1.shape[2] * raw1.shape[3]
                x = ops.instnorm_apply(raw1, ops.instnorm_finalize(part1, V1), ops.ACT_LRELU, inplace=True)
            else:
                x = self._conv_in_act(x, c1)                                     # UpBlock :547-550
            x = self._conv_in_act(x, c2, x1=skip)                                # cat + conv2 :553-554
            self._tap(f"up{i-1}", x)
        logits, _, _ = self._conv(x, P.final, pad=1, out_f32=True)               # :1392
        self._tap("logits", logits)
        if head == "logits":
            return logits
        probs, onehot, labels = ops.head_d2s_softmax(logits, self.dim_output, want_probs=(head == "train"),
                                                     want_onehot=(head == "eval"), want_labels=(head == "labels"))
        if head == "train":
            return probs, mask_list
        return onehot if head == "eval" else labels

    @torch.no_grad()
    def predict_labels(self, x: torch.Tensor) -> torch.Tensor:
        """Eval forward returning the argmax class per voxel as uint8 [B,H,W,D]: the information of
        the reference's one-hot eval output (trans_3DUnet.py:199-201) without materialising it.
        Used by the sliding-window driver.  With CUDA graphs enabled the returned tensor is the
        graph's static output buffer: consume it before the next call with the same input shape."""
        if not x.is_cuda:
            raise RuntimeError("lintransunet_b200.MaskTransUnet runs on CUDA (sm_100a) only; there is no CPU path")
        return self._run(x, "labels")

    @torch.no_grad()
    def forward_logits(self, x: torch.Tensor) -> torch.Tensor:
        """The `decode.final_block` tap as fp32 channels-last [B,H/2,W/2,D,4*dim_output]."""
        out = self._run(x, "logits")
        return out.clone() if self.use_cuda_graphs else out


Model_Dict = {"MaskTransUnet": MaskTransUnet}


def get_model_dict(name: str):
    """model/trans_3DUnet.py:215-222.  Only MaskTransUnet is alive in the reference."""
    if name not in Model_Dict:
        raise KeyError(f"{name}: only MaskTransUnet is provided (the reference's other four registry entries "
                       "crash in the reference itself)")
    return Model_Dict[name]
