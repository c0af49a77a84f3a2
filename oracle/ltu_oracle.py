"""Oracle for the LinTransUNet ``MaskTransUnet`` forward hot path (CPU; device-agnostic, so the GPU tests can also run it
under ``torch.autocast`` on the box to measure the reference algorithm's own bf16 noise floor).

TEST INFRASTRUCTURE ONLY.  This file is a from-scratch, functional restatement
(plain PyTorch on the CPU, fp32 or fp64) of the reference algorithm.  It is the
checker for the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product package ``lintransunet_b200`` never does.

Pinning: the reference is pure Python and importable in the build container, so
the oracle is pinned by ``tools/make_golden.py`` -- it imports the unmodified
reference from /root/reference, feeds it the seeded state_dict produced by
``make_state_dict`` below, and stores the reference outputs under
``tests/golden/``.  ``tests/test_oracle_golden.py`` checks this file against
those vectors (CPU, no GPU needed).

Every function cites the reference lines it restates (paths relative to
/root/reference).  Tensors use the reference layout ``[B, C, H, W, D]``.
The sliding-window part (MONAI 0.7.0, not vendored by the reference, not
installable offline) is in ``oracle/sliding_window.py`` and is "parity
unpinned" -- see its header.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------- config
@dataclass
class UnetConfig:
    """Constructor arguments of MaskTransUnet (model/trans_3DUnet.py:161-162)."""
    num_layers: Sequence[int] = (16, 32, 64, 128, 256)
    roi_size_list: Sequence[int] = (100, 65, 40, 25, 10)
    is_roi_list: Sequence[bool] = (False, True, True, True, True)
    dim_input: int = 1
    dim_output: int = 2
    kernel_size: int = 3
    n_attn_layers: int = 8      # ROIDecoder N (model/Unet_3Dblock.py:1294)
    head_dim: int = 32          # nhead_lens (model/Unet_3Dblock.py:1294)

    @property
    def levels(self) -> int:
        return len(self.num_layers)

    def bridge_dims(self, i: int) -> Tuple[int, int, int]:
        """(in_dim, d_model, nhead) of decode.bridge_list[i] (Unet_3Dblock.py:1311-1326)."""
        if i == self.levels - 1:
            c = self.num_layers[-1]
            return c, c, c // self.head_dim
        c = self.num_layers[i]
        dm = min(4 * c, 256)
        return c, dm, dm // 32

    def roi_consts(self, i: int) -> Dict[str, int]:
        """ROI constants of ROIBridge i (model/Unet_3Dblock.py:697-714)."""
        r = self.roi_size_list[i]
        e = int(1.2 * r)
        ew = int(e * 0.6)
        return dict(h_roi=r, w_roi=int(r * 0.6), eval_h=e, eval_w=ew,
                    min_h=e // 2, min_w=ew // 2)


def state_dict_spec(cfg: UnetConfig) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(key, shape, kind) for every tensor of the reference state_dict, in the
    reference's registration order is NOT required -- only names and shapes are
    interface.  kind in {conv_w, conv_b, lin_w, lin_b, ln_w, ln_b}."""
    L = list(cfg.num_layers)
    k = cfg.kernel_size
    out: List[Tuple[str, Tuple[int, ...], str]] = []

    def conv(name, cout, cin, ks=k):
        out.append((name + ".weight", (cout, cin, ks, ks, ks), "conv_w"))
        out.append((name + ".bias", (cout,), "conv_b"))

    def lin(name, cout, cin):
        out.append((name + ".weight", (cout, cin), "lin_w"))
        out.append((name + ".bias", (cout,), "lin_b"))

    def ln(name, c):
        out.append((name + ".weight", (c,), "ln_w"))
        out.append((name + ".bias", (c,), "ln_b"))

    def layers(prefix, dm):
        for j in range(cfg.n_attn_layers):
            p = f"{prefix}.layers.{j}"
            for q in range(4):
                lin(f"{p}.self_attn.linears.{q}", dm, dm)
            lin(f"{p}.linear1", 2 * dm, dm)
            lin(f"{p}.linear2", dm, 2 * dm)
            ln(f"{p}.layer_norm1", dm)
            ln(f"{p}.layer_norm2", dm)

    conv("encode.input_block", L[0], cfg.dim_input * 4)
    for i in range(1, len(L)):
        conv(f"encode.block_list.{i-1}.conv1", L[i - 1], L[i - 1])
        conv(f"encode.block_list.{i-1}.conv2", L[i], L[i - 1])
    for i in range(len(L) - 1):
        if not cfg.is_roi_list[i]:
            continue
        cin, dm, _ = cfg.bridge_dims(i)
        p = f"decode.bridge_list.{i}.transformer"
        conv(f"{p}.down_embed.module_list.0.0", dm, cin, 3)
        conv(f"{p}.up_embed.module_list.0.1", cin, dm, 3)
        out.append((f"{p}.pos_encoder.proj.weight", (dm, 1, 3, 3, 3), "conv_w"))
        out.append((f"{p}.pos_encoder.proj.bias", (dm,), "conv_b"))
        layers(p, dm)
    ib = len(L) - 1
    _, dm, _ = cfg.bridge_dims(ib)
    p = f"decode.bridge_list.{ib}.transformer"
    for j in range(cfg.n_attn_layers):
        out.append((f"{p}.pos_encoders.{j}.proj.weight", (dm, 1, 3, 3, 3), "conv_w"))
        out.append((f"{p}.pos_encoders.{j}.proj.bias", (dm,), "conv_b"))
    layers(p, dm)
    for i in range(1, len(L)):
        conv(f"decode.mask_conv_list.{i-1}", cfg.dim_output, L[i])
        a = f"decode.att_conv_list.{i-1}"
        conv(f"{a}.W_x.0", L[i - 1], L[i - 1], 1)
        conv(f"{a}.W_g.0", L[i - 1], L[i], 1)
        conv(f"{a}.psi.0", 1, L[i - 1], 1)
    for i in range(1, len(L)):
        conv(f"decode.block_list.{i-1}.conv1", L[-i - 1], L[-i])
        conv(f"decode.block_list.{i-1}.conv2", L[-i - 1], 2 * L[-i - 1])
    conv("decode.final_block", cfg.dim_output * 4, L[0])
    return out


def make_state_dict(cfg: UnetConfig, seed: int = 0, dtype=torch.float32) -> Dict[str, Tensor]:
    """Seeded synthetic weights with the reference's names/shapes.

    Same distribution family as PyTorch's default init (U(-1/sqrt(fan_in), +))
    so activations are realistic; LayerNorm affine is perturbed away from
    (1, 0) so that the affine path is exercised.  Both the reference (via
    load_state_dict) and the product consume this dict, so no RNG stream of
    either implementation has to match.
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    fan_in_of: Dict[str, int] = {}
    for key, shape, kind in state_dict_spec(cfg):
        if kind in ("conv_w", "lin_w"):
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            fan_in_of[key.rsplit(".", 1)[0]] = fan_in
            b = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * b
        elif kind in ("conv_b", "lin_b"):
            b = 1.0 / math.sqrt(fan_in_of[key.rsplit(".", 1)[0]])
            t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * b
        elif kind == "ln_w":
            t = 1.0 + 0.1 * (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1)
        elif kind == "ln_b":
            t = 0.05 * (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1)
        else:  # pragma: no cover
            raise ValueError(kind)
        sd[key] = t.to(dtype)
    return sd


def make_input(shape: Sequence[int], seed: int = 1, blob: bool = False) -> Tensor:
    """Seeded synthetic CT patch ~N(0,1) (z-scored HU, dataset/CT_pancreas_ids.py:150-152).
    blob=True adds a smooth bright ellipsoid so the mask heads see structure."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    x = torch.randn(tuple(shape), generator=g, dtype=torch.float32)
    if blob:
        _, _, H, W, D = shape
        hh = torch.linspace(-1, 1, H).view(H, 1, 1)
        ww = torch.linspace(-1, 1, W).view(1, W, 1)
        dd = torch.linspace(-1, 1, D).view(1, 1, D)
        r2 = ((hh - 0.1) / 0.45) ** 2 + ((ww + 0.15) / 0.35) ** 2 + (dd / 0.8) ** 2
        x = x + 2.5 * torch.exp(-2.0 * r2)
    return x


# ------------------------------------------------------------------- transformer ops
def efficient_attention(q: Tensor, k: Tensor, v: Tensor) -> Tensor:
    """linear_attention, model/trans_block.py:41-67 (mask=None path; the dropout
    only touches the returned score tensor, never `out` -- :62-65).
    q,k,v: [B, h, N, d] -> [B, h, N, d]."""
    d = q.shape[-1]
    qh = torch.softmax(q, dim=-1) / math.sqrt(d)          # :50
    kh = torch.softmax(k, dim=-2)                         # :59  (over the N tokens)
    ctx = torch.matmul(kh.transpose(-1, -2), v)           # :60  [B,h,d,d]
    return torch.matmul(qh, ctx)                          # :65


def multihead_attention(x: Tensor, sd: Dict[str, Tensor], prefix: str, nhead: int) -> Tensor:
    """MultihAttention.forward, model/trans_block.py:148-166. x: [B,N,C]."""
    B, N, C = x.shape
    d = C // nhead
    proj = [F.linear(x, sd[f"{prefix}.linears.{i}.weight"], sd[f"{prefix}.linears.{i}.bias"])
            for i in range(3)]
    q, k, v = [t.view(B, N, nhead, d).transpose(1, 2) for t in proj]   # :155-157
    o = efficient_attention(q, k, v)
    o = o.transpose(1, 2).reshape(B, N, C)                              # :165
    return F.linear(o, sd[f"{prefix}.linears.3.weight"], sd[f"{prefix}.linears.3.bias"])


def encoder_layer(x: Tensor, sd: Dict[str, Tensor], prefix: str, nhead: int) -> Tensor:
    """SelfAttentionLayer.forward (eval / dropout-free), model/trans_block.py:203-211.
    Post-norm, LN eps 1e-6 (:183), erf-GELU (:201), FFN width 2C."""
    C = x.shape[-1]
    a = multihead_attention(x, sd, f"{prefix}.self_attn", nhead)
    x = F.layer_norm(x + a, (C,), sd[f"{prefix}.layer_norm1.weight"],
                     sd[f"{prefix}.layer_norm1.bias"], eps=1e-6)
    f = F.linear(F.gelu(F.linear(x, sd[f"{prefix}.linear1.weight"], sd[f"{prefix}.linear1.bias"])),
                 sd[f"{prefix}.linear2.weight"], sd[f"{prefix}.linear2.bias"])
    return F.layer_norm(x + f, (C,), sd[f"{prefix}.layer_norm2.weight"],
                        sd[f"{prefix}.layer_norm2.bias"], eps=1e-6)


def pos_embedding(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """Conv3dPosEmbedding.forward, model/trans_block.py:86-96, restated in the
    native [B,C,H,W,D] axes: the reference applies the depthwise conv on the
    (D,H,W)-permuted view (model/Unet_3Dblock.py:259-270, :481-490), which equals a
    conv on (H,W,D) with kernel w'[c,0,kh,kw,kd] = w[c,0,kd,kh,kw]."""
    wp = w.permute(0, 1, 3, 4, 2)
    return x + F.conv3d(x, wp, b, stride=1, padding=1, groups=x.shape[1])


def transformer_stack(x: Tensor, sd: Dict[str, Tensor], prefix: str, nhead: int,
                      pos_w: Tensor, pos_b: Tensor, n_layers: int,
                      taps: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """The 8-layer loop shared by PosAttention3DBlock.forward (Unet_3Dblock.py:249-274)
    and EmbedAttention3DBlock.forward (:481-497): tokens are a flattening of the
    volume (order irrelevant: every op but the pos-conv is permutation-equivariant),
    positional conv is applied once, after layer 0."""
    B, C, H, W, D = x.shape
    t = x.permute(0, 2, 3, 4, 1).reshape(B, H * W * D, C)
    for i in range(n_layers):
        t = encoder_layer(t, sd, f"{prefix}.layers.{i}", nhead)
        if taps is not None:
            taps[f"{prefix}.layers.{i}"] = t.reshape(B, H, W, D, C)
        if i == 0:
            vol = t.reshape(B, H, W, D, C).permute(0, 4, 1, 2, 3)
            vol = pos_embedding(vol, pos_w, pos_b)
            t = vol.permute(0, 2, 3, 4, 1).reshape(B, H * W * D, C)
    return t.reshape(B, H, W, D, C).permute(0, 4, 1, 2, 3)


# ------------------------------------------------------------------------ conv stages
def inorm(x: Tensor) -> Tensor:
    """nn.InstanceNorm3d defaults: no affine, eps 1e-5, biased variance."""
    return F.instance_norm(x, eps=1e-5)


def lrelu(x: Tensor) -> Tensor:
    return F.leaky_relu(x, 0.01)


def space_to_depth(img: Tensor, k: int = 2) -> Tensor:
    """windows_embedding, model/Unet_3Dblock.py:123-136: channel = kh*k + kw."""
    B, _, H, W, D = img.shape
    t = img.reshape(B, H // k, k, W // k, k, D)
    return t.permute(0, 2, 4, 1, 3, 5).reshape(B, k * k, H // k, W // k, D)


def depth_to_space(img: Tensor, k: int = 2) -> Tensor:
    """windows_unembedding, model/Unet_3Dblock.py:138-152: in-channel = c*k*k + kh*k + kw."""
    B, C, H, W, D = img.shape
    t = img.reshape(B, C // (k * k), k, k, H, W, D)
    return t.permute(0, 1, 4, 2, 5, 3, 6).reshape(B, C // (k * k), H * k, W * k, D)


def encoder_forward(x: Tensor, sd: Dict[str, Tensor], cfg: UnetConfig,
                    taps: Optional[Dict[str, Tensor]] = None):
    """Encoder.forward (Unet_3Dblock.py:596-607) with DownBlock.forward (:325-341);
    conv2 stride (2,2,1) for blocks 0,2 and (2,2,2) for 1,3 (:584)."""
    x = space_to_depth(x, 2)
    x = lrelu(inorm(F.conv3d(x, sd["encode.input_block.weight"], sd["encode.input_block.bias"],
                             padding=1)))
    skips = []
    for i in range(cfg.levels - 1):
        p = f"encode.block_list.{i}"
        s = lrelu(inorm(F.conv3d(x, sd[f"{p}.conv1.weight"], sd[f"{p}.conv1.bias"], padding=1))) + x
        skips.append(s)
        stride = (2, 2, i % 2 + 1)
        x = lrelu(inorm(F.conv3d(s, sd[f"{p}.conv2.weight"], sd[f"{p}.conv2.bias"],
                                 stride=stride, padding=cfg.kernel_size // 2)))
        if taps is not None:
            taps[f"encode.skip{i}"] = s
    if taps is not None:
        taps["encode.bottle"] = x
    return x, skips


def attention_gate(skip: Tensor, up: Tensor, sd: Dict[str, Tensor], prefix: str) -> Tensor:
    """SpatialAttention3DBlock.forward, Unet_3Dblock.py:217-221: all convs 1x1x1."""
    a = inorm(F.conv3d(skip, sd[f"{prefix}.W_x.0.weight"], sd[f"{prefix}.W_x.0.bias"]))
    b = inorm(F.conv3d(up, sd[f"{prefix}.W_g.0.weight"], sd[f"{prefix}.W_g.0.bias"]))
    return torch.sigmoid(F.conv3d(F.relu(a + b), sd[f"{prefix}.psi.0.weight"],
                                  sd[f"{prefix}.psi.0.bias"]))


# ------------------------------------------------------------------------ ROI plumbing
def _quantile_indices(profile: Tensor) -> Tuple[float, float, float]:
    """get_min_max_indice, Unet_3Dblock.py:37-49.  profile: int64 [S].  fp32 arithmetic
    like the reference (int64/int64 true-divide yields the default float dtype)."""
    S = profile.shape[0]
    tot = int(profile.sum())
    if tot == 0:
        mid = S / 2
        return mid - 1, mid + 1, mid
    profile = profile.cpu()           # the reference does this arithmetic in fp32 on whatever device; indices are exact
    r = torch.cumsum(profile, 0).to(torch.float32) / torch.tensor(float(tot), dtype=torch.float32)
    lo_t = torch.tensor(0.001, dtype=torch.float32)
    hi_t = torch.tensor(1 - 0.001, dtype=torch.float32)
    md_t = torch.tensor(0.5, dtype=torch.float32)
    lo = int((r < lo_t).sum())        # searchsorted left : first i with r[i] >= v
    hi = int((r <= hi_t).sum())       # searchsorted right: first i with r[i] >  v
    mid = int((r <= md_t).sum())
    return float(lo), float(hi), float(mid)


def roi_boxes(fg: Tensor, min_h: int, min_w: int, thr: float = 0.5) -> Tensor:
    """ROIBridge.get_mask_boundary2, Unet_3Dblock.py:821-873 on `fg >= thr` (:738-739).
    fg: [B,1,h,w,d] -> fp32 [B,6] = [x0,y0,0,x1,y1,d-1].  Both clamp tests use the size
    computed BEFORE the first clamp (:847-848), so both may fire."""
    m = fg >= thr
    B = m.shape[0]
    H, W, D = m.shape[-3:]
    prof_h = m.sum(dim=(3, 4)).reshape(B, H).to(torch.int64)
    prof_w = m.sum(dim=(2, 4)).reshape(B, W).to(torch.int64)
    f32 = lambda v: torch.tensor(v, dtype=torch.float32)
    box = torch.zeros(B, 6, dtype=torch.float32)
    for b in range(B):
        for (prof, S, mn, i0, i1) in ((prof_h[b], H, min_h, 0, 3), (prof_w[b], W, min_w, 1, 4)):
            lo, hi, mid = _quantile_indices(prof)
            lo, hi, mid = f32(lo), f32(hi), f32(mid)
            size = hi - lo
            if size < mn:
                lo = torch.maximum(mid - mn / 2, f32(0.0))
                hi = torch.minimum(mid + mn / 2, f32(float(S)))
            if size > (S - mn):
                lo = torch.maximum(mid - (S - mn) / 2, f32(0.0))
                hi = torch.minimum(mid + (S - mn) / 2, f32(float(S)))
            box[b, i0], box[b, i1] = lo, hi
        box[b, 2], box[b, 5] = 0.0, float(D - 1)
    return box.to(fg.device)


def fisheye_forward_coords(x0: Tensor, x1: Tensor, h: int, roi: int, eroi: int) -> Tensor:
    """get_transfer_index, Unet_3Dblock.py:51-64.  x0,x1: fp32 [B,1]; returns the
    normalised grid coordinate [B, eroi].  The two fix-ups are sequential and the second
    test sees the already-updated value."""
    i = torch.arange(0, eroi, dtype=torch.float32, device=x0.device)
    k2 = (x1 - x0) / (roi - 1)
    k1 = (h - x1 + x0) / (eroi - roi)
    t = i * k2 + x0 * (1 - k2 / k1)
    alt = t * (k1 / k2) + x0 * (1 - k1 / k2)
    t = torch.where(t <= x0, alt, t)
    alt = t * (k1 / k2) + x1 * (1 - k1 / k2)
    t = torch.where(t >= x1, alt, t)
    return t * 2.0 / h - 1


def fisheye_back_coords(x0: Tensor, x1: Tensor, h: int, roi: int, eroi: int) -> Tensor:
    """get_transfer_back_index, Unet_3Dblock.py:66-82; normalises by /eroi (not eroi-1, :81)."""
    p = torch.arange(0, h + 1, dtype=torch.float32, device=x0.device)
    k2 = roi / (x1 - x0)
    k1 = (eroi - roi) / (h - x1 + x0)
    p0 = x0 * k1
    p1 = eroi - (h - x1) * k1
    t = p * k2 + p0 * (1 - k2 / k1)
    alt = t * (k1 / k2) + p0 * (1 - k1 / k2)
    t = torch.where(t <= p0, alt, t)
    alt = t * (k1 / k2) + p1 * (1 - k1 / k2)
    t = torch.where(t >= p1, alt, t)
    return t * 2 / eroi - 1


def _axis_taps(coord: Tensor, size: int):
    """Bilinear taps of F.grid_sample(align_corners=True, padding_mode='zeros') along one
    axis: coord [B,n] normalised -> (i0, i1, w0, w1) with out-of-range taps weighted 0."""
    pos = (coord + 1) / 2 * (size - 1)
    f = torch.floor(pos)
    w1 = pos - f
    w0 = 1 - w1
    i0 = f.to(torch.int64)
    i1 = i0 + 1
    w0 = torch.where((i0 >= 0) & (i0 <= size - 1), w0, torch.zeros_like(w0))
    w1 = torch.where((i1 >= 0) & (i1 <= size - 1), w1, torch.zeros_like(w1))
    return i0.clamp(0, size - 1), i1.clamp(0, size - 1), w0, w1


def separable_resample(x: Tensor, ch: Tensor, cw: Tensor) -> Tensor:
    """The 2-D bilinear grid_sample of roi_alignment2 / post_processing2
    (Unet_3Dblock.py:1024-1039, :1104-1117) restated as two 1-D interpolations: the
    sampling grid is an outer product (row coords depend only on i, column coords only
    on j) and is shared by every depth slice and channel.
    x: [B,C,h,w,d]; ch: [B,nh], cw: [B,nw] normalised coords -> [B,C,nh,nw,d]."""
    B, C, h, w, d = x.shape
    ch = ch.to(x.dtype)
    cw = cw.to(x.dtype)
    out = []
    for b in range(B):
        i0, i1, a0, a1 = _axis_taps(ch[b:b + 1], h)
        j0, j1, b0, b1 = _axis_taps(cw[b:b + 1], w)
        xb = x[b]
        rows = xb[:, i0[0]] * a0[0].view(1, -1, 1, 1) + xb[:, i1[0]] * a1[0].view(1, -1, 1, 1)
        o = rows[:, :, j0[0]] * b0[0].view(1, 1, -1, 1) + rows[:, :, j1[0]] * b1[0].view(1, 1, -1, 1)
        out.append(o)
    return torch.stack(out, 0)


def roi_bridge(x: Tensor, fg: Tensor, sd: Dict[str, Tensor], cfg: UnetConfig, i: int,
               forced_box: Optional[Tensor] = None,
               taps: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """ROIBridge.forward, Unet_3Dblock.py:717-755: box -> fisheye resample ->
    EmbedAttention3DBlock (:469-501) -> resample back; the result replaces the skip."""
    rc = cfg.roi_consts(i)
    cin, dm, nhead = cfg.bridge_dims(i)
    B, C, h, w, d = x.shape
    box = (roi_boxes(fg, rc["min_h"], rc["min_w"]) if forced_box is None
           else forced_box.to(device=x.device, dtype=torch.float32))
    if taps is not None:
        taps[f"box{i}"] = box
    x0, y0, x1, y1 = box[:, 0:1], box[:, 1:2], box[:, 3:4], box[:, 4:5]
    ch = fisheye_forward_coords(x0, x1, h - 1, rc["h_roi"], rc["eval_h"])
    cw = fisheye_forward_coords(y0, y1, w - 1, rc["w_roi"], rc["eval_w"])
    roi = separable_resample(x, ch, cw)
    if taps is not None:
        taps[f"roi_in{i}"] = roi
    p = f"decode.bridge_list.{i}.transformer"
    t = lrelu(inorm(F.conv3d(roi, sd[f"{p}.down_embed.module_list.0.0.weight"],
                             sd[f"{p}.down_embed.module_list.0.0.bias"], stride=2, padding=1)))
    t = transformer_stack(t, sd, p, nhead, sd[f"{p}.pos_encoder.proj.weight"],
                          sd[f"{p}.pos_encoder.proj.bias"], cfg.n_attn_layers, taps)
    t = F.interpolate(t, scale_factor=2, mode="nearest")          # nn.Upsample(scale_factor=2) :421
    t = lrelu(inorm(F.conv3d(t, sd[f"{p}.up_embed.module_list.0.1.weight"],
                             sd[f"{p}.up_embed.module_list.0.1.bias"], padding=1)))
    if taps is not None:
        taps[f"roi_out{i}"] = t
    bh = fisheye_back_coords(x0, x1, h - 1, rc["h_roi"], rc["eval_h"])
    bw = fisheye_back_coords(y0, y1, w - 1, rc["w_roi"], rc["eval_w"])
    return separable_resample(t, bh, bw)


def trilinear_up(x: Tensor, scale: Tuple[int, int, int]) -> Tensor:
    """nn.Upsample(mode='trilinear', align_corners=True), Unet_3Dblock.py:1341-1345."""
    return F.interpolate(x, scale_factor=tuple(float(s) for s in scale), mode="trilinear",
                         align_corners=True)


def decoder_forward(bottle: Tensor, skips: List[Tensor], sd: Dict[str, Tensor], cfg: UnetConfig,
                    forced_boxes: Optional[Dict[int, Tensor]] = None,
                    taps: Optional[Dict[str, Tensor]] = None):
    """ROIDecoder.forward, Unet_3Dblock.py:1359-1396."""
    n = cfg.levels
    ib = n - 1
    p = f"decode.bridge_list.{ib}.transformer"
    _, dm, nhead = cfg.bridge_dims(ib)
    x = transformer_stack(bottle, sd, p, nhead, sd[f"{p}.pos_encoders.0.proj.weight"],
                          sd[f"{p}.pos_encoders.0.proj.bias"], cfg.n_attn_layers, taps)
    if taps is not None:
        taps["bridge_bottle"] = x
    mask_list = []
    for i in range(1, n):
        scale = (2, 2, 2) if (n - i) % 2 == 0 else (2, 2, 1)        # :1375-1378
        x = trilinear_up(x, scale)
        mk = n - 1 - i                                               # index of *_list[-i]
        logits = F.conv3d(x, sd[f"decode.mask_conv_list.{mk}.weight"],
                          sd[f"decode.mask_conv_list.{mk}.bias"], padding=1)
        mask = torch.softmax(logits, dim=1)
        mask_list.append(mask)
        skip = skips[-i]
        gate = attention_gate(skip, x, sd, f"decode.att_conv_list.{mk}")
        skip = skip * gate
        fg = (1 - mask[:, 0]).unsqueeze(1)
        bi = n - 1 - i                                               # bridge_list[-i-1]
        if cfg.is_roi_list[bi]:
            fb = None if forced_boxes is None else forced_boxes.get(bi)
            skip = roi_bridge(skip, fg, sd, cfg, bi, fb, taps)
        if taps is not None:
            taps[f"bridged_skip{bi}"] = skip
        q = f"decode.block_list.{i-1}"
        x = lrelu(inorm(F.conv3d(x, sd[f"{q}.conv1.weight"], sd[f"{q}.conv1.bias"], padding=1)))
        x = torch.cat((x, skip), dim=1)
        x = lrelu(inorm(F.conv3d(x, sd[f"{q}.conv2.weight"], sd[f"{q}.conv2.bias"], padding=1)))
        if taps is not None:
            taps[f"decode.up{i-1}"] = x
    logits = F.conv3d(x, sd["decode.final_block.weight"], sd["decode.final_block.bias"], padding=1)
    probs = torch.softmax(depth_to_space(logits, 2), dim=1)
    return logits, probs, mask_list


def mask_trans_unet_forward(x: Tensor, sd: Dict[str, Tensor], cfg: UnetConfig,
                            forced_boxes: Optional[Dict[int, Tensor]] = None,
                            want_taps: bool = False) -> Dict[str, object]:
    """MaskTransUnet.forward, model/trans_3DUnet.py:181-204, dropout-free.
    Returns logits (the decode.final_block tap), probs, mask_list, onehot (eval output)
    and the ROI boxes / intermediate taps."""
    assert cfg.dim_input == 1, "windows_embedding requires a single input channel (Unet_3Dblock.py:132)"
    sd = {k: v.to(device=x.device, dtype=x.dtype) for k, v in sd.items()}
    taps: Optional[Dict[str, Tensor]] = {} if want_taps else None
    box_taps: Dict[str, Tensor] = {} if taps is None else taps
    bottle, skips = encoder_forward(x, sd, cfg, taps)
    logits, probs, mask_list = decoder_forward(bottle, skips, sd, cfg, forced_boxes, box_taps)
    idx = torch.argmax(probs, dim=1, keepdim=True)
    onehot = torch.zeros_like(probs).scatter_(1, idx, 1)             # trans_3DUnet.py:199-201
    boxes = {int(k[3:]): v for k, v in box_taps.items() if k.startswith("box")}
    return dict(logits=logits, probs=probs, mask_list=mask_list, onehot=onehot,
                boxes=boxes, taps=taps)
