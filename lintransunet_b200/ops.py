"""Tensor-level wrappers of the C ABI (include/ltu_b200.h).

Every function takes CUDA tensors in the channels-last layout ``[B, H, W, D, C]`` (or token
matrices ``[B, N, C]``), allocates outputs with torch (the library never allocates) and
launches on the current stream of the tensor's device.  Nothing here computes on the CPU and
nothing falls back to PyTorch kernels: a CPU tensor raises.
"""
from __future__ import annotations

import os
from ctypes import c_void_p
from typing import Optional, Tuple

import torch

from . import _native
from ._native import check

F32, BF16 = 0, 1
ACT_NONE, ACT_LRELU = 0, 1
USE_HALO_CONV = os.environ.get("LTU_DISABLE_HALO", "0") != "1"     # A/B switch for the small-channel conv kernel
USE_TC3_CONV = os.environ.get("LTU_DISABLE_TC3", "0") != "1"       # A/B switch for the TMA-halo tcgen05 conv kernel
USE_CONCAT_TC3 = os.environ.get("LTU_CONCAT_TC3", "1") == "1"      # 32 + 32 input channels: concatenate once, run on the TMA-halo kernel


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported activation dtype {t.dtype} (float32 or bfloat16)")


def _p(t: Optional[torch.Tensor]):
    return None if t is None else c_void_p(t.data_ptr())


def _chk(*ts: Optional[torch.Tensor]) -> torch.device:
    dev = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("lintransunet_b200 ops need CUDA tensors: there is no CPU fallback")
        if not t.is_contiguous():
            raise RuntimeError("lintransunet_b200 ops need contiguous tensors")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"tensors on different devices: {t.device} vs {dev}")
    return dev


class KernelProfiler:
    """Per-kernel device timing with CUDA events on the launching stream (bench.py roofline):
    every profiled launch records (name, start, end, algorithmic bytes, algorithmic flops)."""

    def __init__(self):
        self.records = []

    def summary(self):
        """name -> dict(launches, ms, bytes, flops, flops_exec); call after torch.cuda.synchronize().  `flops` is the
        algorithmic count (un-folded), `flops_exec` what the kernel executes (8/27 of it for the folded
        nearest-x2 + 3x3x3 convolutions)."""
        out = {}
        for name, e0, e1, nbytes, flops, fexec in self.records:
            d = out.setdefault(name, dict(launches=0, ms=0.0, bytes=0, flops=0, flops_exec=0))
            d["launches"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["bytes"] += nbytes
            d["flops"] += flops
            d["flops_exec"] += fexec
        return out


_profiler: Optional[KernelProfiler] = None


def set_profiler(p: Optional[KernelProfiler]) -> None:
    global _profiler
    _profiler = p


class _Guard:
    """Make `dev` current for the launch (DataParallel replicas run on foreign devices) and, when a
    KernelProfiler is installed, bracket the launch with CUDA events."""
    __slots__ = ("dev", "prev", "prof", "e0")

    def __init__(self, dev: torch.device, prof=None):
        self.dev = dev
        self.prev = None
        self.prof = prof if _profiler is not None else None
        self.e0 = None

    def __enter__(self):
        cur = torch.cuda.current_device()
        if self.dev.index is not None and self.dev.index != cur:
            self.prev = cur
            torch.cuda.set_device(self.dev)
        if self.prof is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return c_void_p(torch.cuda.current_stream().cuda_stream)

    def __exit__(self, *exc):
        if self.prof is not None and _profiler is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _profiler.records.append((self.prof[0], self.e0, e1, self.prof[1], self.prof[2],
                                      self.prof[3] if len(self.prof) > 3 else self.prof[2]))
        if self.prev is not None:
            torch.cuda.set_device(self.prev)
        return False


# ------------------------------------------------------------------ a1: attention core
def kv_reduce(k: torch.Tensor, v: torch.Tensor, heads: int, w_o: Optional[torch.Tensor] = None):
    """k, v: [B, N, C] views with a common row stride (e.g. slices of a fused QKV buffer).
    Returns ctx fp32 [B, heads, 32, 32] (model/trans_block.py:59-60).  With w_o (the output projection's bf16 weight
    [C, C]; bf16 k / v, 4 or 8 heads) the merge kernel also writes the per-sample weight of ctx_project and the result is
    (ctx, W_b)."""
    B, N, C = k.shape
    assert C == heads * 32 and v.shape == k.shape
    assert k.stride(2) == 1 and v.stride(2) == 1 and k.stride(1) == v.stride(1)
    assert k.stride(0) == N * k.stride(1) and v.stride(0) == N * v.stride(1)
    if not (k.is_cuda and v.is_cuda):
        raise RuntimeError("lintransunet_b200 ops need CUDA tensors: there is no CPU fallback")
    L = _native.lib()
    with _Guard(k.device, ("kv_reduce", 2 * B * N * C * k.element_size(), 2 * B * N * C * 32)) as st:
        ws_bytes = L.ltu_kv_reduce_workspace(B, N, heads)
        ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=k.device)
        ctx = torch.empty(B, heads, 32, 32, dtype=torch.float32, device=k.device)
        if w_o is not None:
            if k.dtype != torch.bfloat16 or w_o.dtype != torch.bfloat16 or tuple(w_o.shape) != (C, C) or not w_o.is_contiguous():
                raise TypeError("kv_reduce(w_o=...) needs bf16 k / v and the contiguous bf16 [C, C] weight")
            wb = torch.empty(B, C, C, dtype=torch.bfloat16, device=k.device)
            check(L.ltu_kv_reduce_project(_p(k), _p(v), k.stride(1), _p(ctx), _p(ws), ws_bytes, B, N, heads, _p(w_o), _p(wb), st),
                  "ltu_kv_reduce_project")
            return ctx, wb
        check(L.ltu_kv_reduce(_p(k), _p(v), k.stride(1), _p(ctx), _p(ws), ws_bytes, B, N, heads, _dt(k), st),
              "ltu_kv_reduce")
    return ctx


def q_readout(q: torch.Tensor, ctx: torch.Tensor, heads: int) -> torch.Tensor:
    """q: [B, N, C] view (row stride free), ctx fp32 [B,heads,32,32] -> out [B, N, C]
    (model/trans_block.py:50,:65,:165)."""
    B, N, C = q.shape
    assert C == heads * 32 and q.stride(2) == 1 and q.stride(0) == N * q.stride(1)
    _chk(ctx)
    if not q.is_cuda:
        raise RuntimeError("lintransunet_b200 ops need CUDA tensors: there is no CPU fallback")
    out = torch.empty(B, N, C, dtype=q.dtype, device=q.device)
    with _Guard(q.device, ("q_readout", 2 * B * N * C * q.element_size(), 2 * B * N * C * 32)) as st:
        check(_native.lib().ltu_q_readout(_p(q), q.stride(1), _p(ctx), _p(out), C, B, N, heads, _dt(q), st),
              "ltu_q_readout")
    return out


def linear_attention_bwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, ctx: torch.Tensor, dout: torch.Tensor,
                         heads: int):
    """Backward of the attention core (model/trans_block.py:41-67): q, k, v, dout [B, N, C] views (k and v with a common
    row stride), ctx fp32 [B,heads,32,32] from kv_reduce -> dqkv [B, N, 3C] (dq | dk | dv, activation dtype): one
    buffer, so the backward of a fused QKV projection is a single GEMM."""
    B, N, C = q.shape
    assert C == heads * 32 and k.shape == q.shape and v.shape == q.shape and dout.shape == q.shape
    for t in (q, k, v, dout):
        if not t.is_cuda:
            raise RuntimeError("lintransunet_b200 ops need CUDA tensors: there is no CPU fallback")
        assert t.stride(2) == 1 and t.stride(0) == N * t.stride(1) and t.dtype == q.dtype
    assert k.stride(1) == v.stride(1)
    dev = _chk(ctx)
    L = _native.lib()
    ws_bytes = L.ltu_attn_bwd_workspace(B, N, heads)
    ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=dev)
    dqkv = torch.empty(B, N, 3 * C, dtype=q.dtype, device=dev)
    dctx = torch.empty(B, heads, 32, 32, dtype=torch.float32, device=dev)
    kst = torch.empty(B, heads, 3, 32, dtype=torch.float32, device=dev)
    es = q.element_size()
    with _Guard(dev, ("attn_bwd", 10 * B * N * C * es, 10 * B * N * C * 32)) as st:
        check(L.ltu_attn_bwd(_p(q), q.stride(1), _p(k), _p(v), k.stride(1), _p(dout), dout.stride(1), _p(ctx),
                             c_void_p(dqkv.data_ptr()), c_void_p(dqkv.data_ptr() + C * es),
                             c_void_p(dqkv.data_ptr() + 2 * C * es), 3 * C, _p(dctx), _p(kst), _p(ws), ws_bytes,
                             B, N, heads, _dt(q), st), "ltu_attn_bwd")
    return dqkv


def add_layernorm(x: torch.Tensor, res: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
                  eps: float = 1e-6) -> torch.Tensor:
    """LayerNorm(x + res) over the last dim (model/trans_block.py:205-206, :209-210)."""
    dev = _chk(x, res, gamma, beta)
    C = x.shape[-1]
    rows = x.numel() // C
    y = torch.empty_like(x)
    with _Guard(dev, ("add_layernorm", 3 * x.numel() * x.element_size(), 0)) as st:
        check(_native.lib().ltu_add_layernorm(_p(x), _p(res), _p(gamma), _p(beta), _p(y), rows, C, eps, _dt(x), st),
              "ltu_add_layernorm")
    return y


def add_layernorm_split(x_hi: torch.Tensor, x_lo: Optional[torch.Tensor], res: torch.Tensor, gamma: torch.Tensor,
                        beta: torch.Tensor, eps: float = 1e-6):
    """LayerNorm(x_hi + x_lo + res) -> (y_hi, y_lo), all bf16: the split token stream of the bf16 path (the Linear
    layers read y_hi, the next residual add reads y_hi + y_lo; include/ltu_b200.h)."""
    dev = _chk(x_hi, x_lo, res, gamma, beta)
    if x_hi.dtype != torch.bfloat16:
        raise TypeError("add_layernorm_split: bf16 tokens only")
    C = x_hi.shape[-1]
    rows = x_hi.numel() // C
    y_hi, y_lo = torch.empty_like(x_hi), torch.empty_like(x_hi)
    n_units = 4 + (x_lo is not None)
    with _Guard(dev, ("add_layernorm", n_units * x_hi.numel() * 2, 0)) as st:
        check(_native.lib().ltu_add_layernorm_split(_p(x_hi), _p(x_lo), _p(res), _p(gamma), _p(beta), _p(y_hi), _p(y_lo),
                                                    rows, C, eps, st), "ltu_add_layernorm_split")
    return y_hi, y_lo


def add_layernorm_bwd(x: torch.Tensor, res: torch.Tensor, dy: torch.Tensor, gamma: torch.Tensor, eps: float = 1e-6):
    """Backward of LayerNorm(x + res): returns (dz, dgamma, dbeta); dz is the gradient of x AND of res."""
    dev = _chk(x, res, dy, gamma)
    C = x.shape[-1]
    rows = x.numel() // C
    L = _native.lib()
    nbytes = L.ltu_add_layernorm_bwd_workspace(rows, C)
    ws = torch.empty(max(nbytes // 4, 1), dtype=torch.float32, device=dev)
    dz = torch.empty_like(x)
    dgamma = torch.empty(C, dtype=torch.float32, device=dev)
    dbeta = torch.empty(C, dtype=torch.float32, device=dev)
    with _Guard(dev, ("add_layernorm_bwd", 4 * x.numel() * x.element_size(), 0)) as st:
        check(L.ltu_add_layernorm_bwd(_p(x), _p(res), _p(dy), _p(gamma), _p(dz), _p(dgamma), _p(dbeta), _p(ws), nbytes,
                                      rows, C, eps, _dt(x), st), "ltu_add_layernorm_bwd")
    return dz, dgamma, dbeta


def gelu_bwd(x: torch.Tensor, dy: torch.Tensor) -> torch.Tensor:
    """dx = dy * gelu'(x) for the exact-erf GELU; x is the pre-activation."""
    dev = _chk(x, dy)
    dx = torch.empty_like(x)
    with _Guard(dev, ("gelu_bwd", 3 * x.numel() * x.element_size(), 0)) as st:
        check(_native.lib().ltu_gelu_bwd(_p(x), _p(dy), _p(dx), x.numel(), _dt(x), st), "ltu_gelu_bwd")
    return dx


def gelu(x: torch.Tensor) -> torch.Tensor:
    """Out-of-place exact-erf GELU (the training forward keeps the pre-activation for gelu_bwd)."""
    return gelu_(x.clone())


def gelu_(x: torch.Tensor) -> torch.Tensor:
    """In-place exact-erf GELU (model/trans_block.py:201,:208)."""
    dev = _chk(x)
    with _Guard(dev, ("gelu", 2 * x.numel() * x.element_size(), 0)) as st:
        check(_native.lib().ltu_gelu(_p(x), x.numel(), _dt(x), st), "ltu_gelu")
    return x


def loss_sums(p: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """p fp32 [N, C, ...] (probabilities in the reference layout), labels uint8 [N, ...] -> fp32 [N, C, 4]:
    (sum p, sum onehot, sum p*onehot, sum -(1-p)*onehot*log(max(p,1e-6))) per sample and class (include/ltu_b200.h)."""
    dev = _chk(p, labels)
    if p.dtype != torch.float32 or labels.dtype != torch.uint8:
        raise TypeError("loss_sums: fp32 probabilities and uint8 labels")
    N, C = p.shape[0], p.shape[1]
    V = p.numel() // (N * C)
    if labels.numel() != N * V:
        raise ValueError(f"loss_sums: labels {tuple(labels.shape)} do not match probabilities {tuple(p.shape)}")
    L = _native.lib()
    nbytes = L.ltu_loss_sums_workspace(N, C, V)
    ws = torch.empty(max(nbytes // 8, 1), dtype=torch.float64, device=dev)
    sums = torch.empty(N, C, 4, dtype=torch.float32, device=dev)
    with _Guard(dev, ("loss_sums", p.numel() * 4 + labels.numel() * C, 0)) as st:
        check(L.ltu_loss_sums(_p(p), _p(labels), _p(sums), _p(ws), nbytes, N, C, V, st), "ltu_loss_sums")
    return sums


def loss_sums_bwd(p: torch.Tensor, labels: torch.Tensor, gsums: torch.Tensor) -> torch.Tensor:
    """d(loss)/dp (fp32, shape of p) from g = d(loss)/d(loss_sums(p, labels)) fp32 [N, C, 4]."""
    dev = _chk(p, labels, gsums)
    N, C = p.shape[0], p.shape[1]
    V = p.numel() // (N * C)
    dp = torch.empty_like(p)
    with _Guard(dev, ("loss_sums_bwd", 2 * p.numel() * 4 + labels.numel() * C, 0)) as st:
        check(_native.lib().ltu_loss_sums_bwd(_p(p), _p(labels), _p(gsums), _p(dp), N, C, V, st), "ltu_loss_sums_bwd")
    return dp


def label_pool(labels: torch.Tensor, kernel: Tuple[int, int, int]) -> torch.Tensor:
    """F.max_pool3d(labels, kernel_size = stride = kernel) on uint8 labels [N, H, W, D] (the reference's label pyramid)."""
    dev = _chk(labels)
    if labels.dtype != torch.uint8 or labels.dim() != 4:
        raise TypeError("label_pool: uint8 labels [N, H, W, D]")
    N, H, W, D = labels.shape
    kh, kw, kd = kernel
    out = torch.empty(N, H // kh, W // kw, D // kd, dtype=torch.uint8, device=dev)
    with _Guard(dev) as st:
        check(_native.lib().ltu_label_pool(_p(labels), _p(out), N, H, W, D, kh, kw, kd, st), "ltu_label_pool")
    return out


def dropout(x: torch.Tensor, p: float, seed: int, offset: int, channelwise: bool = False, inplace: bool = False) -> torch.Tensor:
    """Training-mode nn.Dropout (or nn.Dropout3d: `channelwise`, one draw per sample and channel) on a channels-last
    tensor [B, ..., C]: keep decisions are a pure function of (seed, offset, index), so the backward is the same call on
    the gradient.  Consumes `dropout_counters(x, channelwise)` Philox counters from `offset` (include/ltu_b200.h)."""
    dev = _chk(x)
    y = x if inplace else torch.empty_like(x)
    C = x.shape[-1]
    per_sample = x.numel() // x.shape[0]
    with _Guard(dev, ("dropout", 2 * x.numel() * x.element_size(), 0)) as st:
        check(_native.lib().ltu_dropout(_p(x), _p(y), x.numel(), C, per_sample, float(p), int(seed) & (2 ** 64 - 1),
                                        int(offset) & (2 ** 64 - 1), int(channelwise), _dt(x), st), "ltu_dropout")
    return y


def dropout_counters(x: torch.Tensor, channelwise: bool = False) -> int:
    n = x.shape[0] * x.shape[-1] if channelwise else x.numel()
    return (n + 3) // 4


def posenc_dwconv3(x: torch.Tensor, w27c: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """x + depthwise 3x3x3 conv(x) + bias on [B,H,W,D,C] (model/trans_block.py:86-96)."""
    dev = _chk(x, w27c, bias)
    B, H, W, D, C = x.shape
    y = torch.empty_like(x)
    with _Guard(dev, ("posenc", 2 * x.numel() * x.element_size(), 2 * 27 * x.numel())) as st:
        check(_native.lib().ltu_posenc_dwconv3(_p(x), _p(w27c), _p(bias), _p(y), B, H, W, D, C, _dt(x), st),
              "ltu_posenc_dwconv3")
    return y


def posenc_dwconv3_split(x_hi: torch.Tensor, x_lo: Optional[torch.Tensor], w27c: torch.Tensor, bias: torch.Tensor):
    """posenc_dwconv3 on the split bf16 token stream: (x_hi, x_lo) [B,H,W,D,C] -> (y_hi, y_lo)."""
    dev = _chk(x_hi, x_lo, w27c, bias)
    if x_hi.dtype != torch.bfloat16:
        raise TypeError("posenc_dwconv3_split: bf16 tokens only")
    B, H, W, D, C = x_hi.shape
    y_hi, y_lo = torch.empty_like(x_hi), torch.empty_like(x_hi)
    with _Guard(dev, ("posenc", (4 if x_lo is not None else 3) * x_hi.numel() * 2, 2 * 27 * x_hi.numel())) as st:
        check(_native.lib().ltu_posenc_dwconv3_split(_p(x_hi), _p(x_lo), _p(w27c), _p(bias), _p(y_hi), _p(y_lo),
                                                     B, H, W, D, C, st), "ltu_posenc_dwconv3_split")
    return y_hi, y_lo


def posenc_dwconv3_bwd(x: torch.Tensor, dy: torch.Tensor, w27c: torch.Tensor):
    """Backward of posenc_dwconv3: returns (dx, dw27c fp32 [27,C], dbias fp32 [C]).  dx is the forward kernel applied
    to dy with the taps reversed (a stride-1 'same' depthwise correlation is its own transpose up to the flip)."""
    dev = _chk(x, dy, w27c)
    B, H, W, D, C = x.shape
    dx = posenc_dwconv3(dy, w27c.flip(0).contiguous(), torch.zeros(C, dtype=torch.float32, device=dev))
    L = _native.lib()
    nbytes = L.ltu_posenc_wgrad_workspace(B, H, W, D, C)
    ws = torch.empty(max(nbytes // 4, 1), dtype=torch.float32, device=dev)
    dw = torch.empty(27, C, dtype=torch.float32, device=dev)
    db = torch.empty(C, dtype=torch.float32, device=dev)
    with _Guard(dev) as st:
        check(L.ltu_posenc_wgrad(_p(x), _p(dy), _p(dw), _p(db), _p(ws), nbytes, B, H, W, D, C, _dt(x), st),
              "ltu_posenc_wgrad")
    return dx, dw, db


# ------------------------------------------------------------------ convolution + InstanceNorm
class SVPack:
    """Super-voxel form of a small-channel stride-1 3x3x3 convolution (include/ltu_b200.h, ltu_conv3d_tc3_masked):
    g = 64 / ci consecutive voxels along D are one 64-channel row.  `w` bf16 [rows32][27 * 64 * n_inputs] is the
    block-Toeplitz repacking, rows ordered (delta, co) for the n_main bf16 channels then (delta, a) for the n_aux fp32
    channels; `mask` 27 bytes: which 16-channel K blocks of a tap are not identically zero."""
    __slots__ = ("w", "bias", "mask", "g", "n_main", "n_aux", "ci", "n_inputs", "blocks")

    def __init__(self, w, bias, mask, g, n_main, n_aux, ci, n_inputs):
        self.w, self.bias, self.mask, self.g = w, bias, mask, g
        self.n_main, self.n_aux, self.ci, self.n_inputs = n_main, n_aux, ci, n_inputs
        self.blocks = sum(bin(b).count("1") for b in mask)          # issued K blocks per 128 rows and input


def sv_pack(w: torch.Tensor, bias: Optional[torch.Tensor], n_main: int, n_aux: int, ci: int, n_inputs: int) -> SVPack:
    """w fp32 [n_main + n_aux, ci * n_inputs, 3, 3, 3] (nn.Conv3d layout, kernel axes (H, W, D)) -> SVPack."""
    g = 64 // ci
    cout = n_main + n_aux
    assert w.shape[0] == cout and w.shape[1] == ci * n_inputs and g * ci == 64
    dev = w.device
    # sel[kg, p, delta, kd] = 1 iff input voxel p of super-voxel (G + kg - 1) is tap kd of output voxel delta of super-voxel G
    sel = torch.zeros(3, g, g, 3, device=dev)
    for kg in range(3):
        for p_ in range(g):
            for d in range(g):
                kd = g * (kg - 1) + p_ - d + 1
                if 0 <= kd <= 2:
                    sel[kg, p_, d, kd] = 1.0
    w6 = w.reshape(cout, n_inputs, ci, 3, 3, 3)                                      # [o, j, c, kh, kw, kd]
    wp = torch.einsum("gpdk,ojchwk->dohwgjpc", sel, w6)                               # [delta, o, kh, kw, kg, j, p, c]
    K = 27 * 64 * n_inputs
    rows = g * cout
    rows32 = (rows + 31) // 32 * 32
    out = torch.zeros(rows32, K, dtype=torch.bfloat16, device=dev)
    main = wp[:, :n_main].reshape(g * n_main, K)
    aux = wp[:, n_main:].reshape(g * n_aux, K)
    out[:g * n_main] = main.to(torch.bfloat16)
    out[g * n_main:rows] = aux.to(torch.bfloat16)
    b = torch.zeros(rows32, dtype=torch.float32, device=dev)
    if bias is not None:
        b[:g * n_main] = bias[:n_main].float().repeat(g)
        b[g * n_main:rows] = bias[n_main:].float().repeat(g)
    mask = []
    for t in range(27):
        kg = t % 3
        m = 0
        for ks in range(4):
            ps = range((16 * ks) // ci, (16 * ks + 15) // ci + 1)
            if any(float(sel[kg, p_].sum()) > 0 for p_ in ps):
                m |= 1 << ks
        mask.append(m)
    return SVPack(out.contiguous(), b, bytes(mask), g, n_main, n_aux, ci, n_inputs)


# Super-voxel form: measured SLOWER than conv3d_halo inside the step (r2_bench3: 54 ms for 133 launches vs 47 ms for the 190
# launches of the mma.sync kernel; the block-Toeplitz weights execute 3-4x the useful flops) -> opt-in, LTU_SV=1
USE_SV_CONV = os.environ.get("LTU_SV", "0") == "1"


def _conv3d_sv(x0, x1, sv: SVPack, want_stats: bool):
    """Run a small-channel 3x3x3 stride-1 convolution in super-voxel form on the TMA-halo tcgen05 kernel."""
    L = _native.lib()
    dev = x0.device
    B, H, W, D, ci = x0.shape
    g = sv.g
    Dg = D // g
    x0v = x0.view(B, H, W, Dg, 64)
    x1v = None if x1 is None else x1.view(B, H, W, Dg, 64)
    cm, ca = g * sv.n_main, g * sv.n_aux
    out = torch.empty(B, H, W, D, sv.n_main, dtype=torch.bfloat16, device=dev) if sv.n_main else None
    aux = torch.empty(B, H, W, D, sv.n_aux, dtype=torch.float32, device=dev) if sv.n_aux else None
    tiles = L.ltu_conv3d_tc3_tiles(B, H, W, Dg, cm, ca, 0)
    partials = torch.empty(B, tiles, cm, 2, dtype=torch.float32, device=dev) if (want_stats and sv.n_main) else None
    V = H * W * D
    cin = ci * sv.n_inputs
    nbytes = (x0.numel() + (0 if x1 is None else x1.numel())) * 2 + (0 if out is None else out.numel() * 2) + \
        (0 if aux is None else aux.numel() * 4)
    flops = 2 * 27 * cin * (sv.n_main + sv.n_aux) * B * V
    rows = B * H * W * Dg
    fexec = 2 * rows * sv.w.shape[0] * 16 * sv.blocks * sv.n_inputs
    with _Guard(dev, ("conv3d_sv", nbytes, flops, fexec)) as st:
        check(L.ltu_conv3d_tc3_masked(_p(x0v), 64, _p(x1v), 0 if x1 is None else 64, B, H, W, Dg, _p(sv.w), sv.w.shape[0],
                                      sv.w.shape[1], _p(sv.bias), cm, _p(out if out is not None else aux), _p(partials), ca,
                                      _p(aux), sv.mask, st), "ltu_conv3d_tc3_masked")
    if partials is not None:
        partials = partials.view(B, tiles * g, sv.n_main, 2)     # columns are (delta, co): delta joins the tile index
        tiles = tiles * g
    return out, partials, tiles, aux


def conv_out_size(n: int, k: int, s: int, pad: int) -> int:
    return (n + 2 * pad - k) // s + 1


def concat2(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """torch.cat([a, b], dim=-1) of two channels-last tensors (model/Unet_3Dblock.py:553), one native copy kernel."""
    dev = _chk(a, b)
    if a.shape[:-1] != b.shape[:-1] or a.dtype != b.dtype:
        raise ValueError("concat2: inputs must agree in everything but the channel count")
    out = torch.empty(*a.shape[:-1], a.shape[-1] + b.shape[-1], dtype=a.dtype, device=dev)
    rows = a.numel() // a.shape[-1]
    with _Guard(dev, ("concat", 2 * out.numel() * out.element_size(), 0)) as st:
        check(_native.lib().ltu_concat2(_p(a), a.shape[-1], _p(b), b.shape[-1], _p(out), rows, _dt(a), st), "ltu_concat2")
    return out


def conv3d(x0: torch.Tensor, w_packed: torch.Tensor, bias: Optional[torch.Tensor], cout: int, ksize: int,
           stride: Tuple[int, int, int] = (1, 1, 1), pad: int = 1, x1: Optional[torch.Tensor] = None,
           up2: bool = False, out_f32: bool = False, want_stats: bool = False,
           w_tc: Optional[torch.Tensor] = None, w_tc_fold: Optional[torch.Tensor] = None, n_aux: int = 0,
           sv: Optional[SVPack] = None):
    """nn.Conv3d on channels-last input(s).  Returns (out, partials, tiles); partials is None
    unless want_stats.  When `w_tc` (bf16 [Cout16][Kpad], see ltu_conv3d_tc) is given and the shape
    qualifies the tcgen05 implicit-GEMM kernel is used, otherwise the CUDA-core kernel."""
    if up2:
        w_tc = w_tc_fold                      # the tensor-core path of an up2 conv needs the folded weights
    dev = _chk(x0, x1, w_packed, bias, w_tc)
    if (sv is not None and USE_SV_CONV and USE_TC3_CONV and x0.dtype == torch.bfloat16 and not up2 and ksize == 3 and pad == 1
            and tuple(stride) == (1, 1, 1) and x0.shape[-1] == sv.ci and x0.shape[3] % sv.g == 0
            and (x1 is not None) == (sv.n_inputs == 2) and (x1 is None or x1.shape == x0.shape)
            and out_f32 == (sv.n_main == 0) and n_aux == (0 if out_f32 else sv.n_aux)):
        out, partials, tiles, aux = _conv3d_sv(x0, x1, sv, want_stats)
        if out_f32:
            return aux, None, tiles
        return (out, partials, tiles, aux) if n_aux else (out, partials, tiles)
    L = _native.lib()
    B, Hi, Wi, Di, C0 = x0.shape
    C1 = 0 if x1 is None else x1.shape[-1]
    if x1 is not None:
        assert x1.shape[:4] == x0.shape[:4] and x1.dtype == x0.dtype
    if (x1 is not None and USE_CONCAT_TC3 and USE_TC3_CONV and w_tc is not None and x0.dtype == torch.bfloat16 and C0 % 64 != 0
            and (C0 + C1) % 64 == 0 and w_tc.shape[-2] % 32 == 0 and (out_f32 or cout % 8 == 0)
            and L.ltu_conv3d_tc_supported(C0, C1, cout + n_aux, ksize, pad) == 1
            and L.ltu_conv3d_tc3_supported(C0 + C1, 0, cout, ksize, stride[0], stride[1], stride[2], pad, int(up2), int(out_f32), n_aux) == 1):
        # two 32-channel inputs: written side by side once, the layer runs as ONE 64-channel input of the TMA-halo tcgen05
        # kernel (its packed weight already has the inputs' channels in this order) instead of the im2col gather kernel
        x0, x1, C0, C1 = concat2(x0, x1), None, C0 + C1, 0
    He, We, De = (2 * Hi, 2 * Wi, 2 * Di) if up2 else (Hi, Wi, Di)
    Ho, Wo, Do = (conv_out_size(He, ksize, stride[0], pad), conv_out_size(We, ksize, stride[1], pad),
                  conv_out_size(De, ksize, stride[2], pad))
    V = Ho * Wo * Do
    ctot = cout + n_aux        # with a fused fp32 head the packed weight/bias carry n_aux extra output rows
    use_tc = (w_tc is not None and x0.dtype == torch.bfloat16 and (out_f32 or cout % 8 == 0)
              and L.ltu_conv3d_tc_supported(C0, C1, ctot, ksize, pad) == 1)
    # small-channel stride-1 layers: shared-memory halo + mma.sync kernel (needs the same bf16 packing)
    use_halo = (USE_HALO_CONV and w_tc is not None and not up2 and x0.dtype == torch.bfloat16
                and (out_f32 or cout % 2 == 0) and w_tc.shape[0] >= (32 if ctot > 16 else 16)
                and L.ltu_conv3d_halo_supported(C0, C1, ctot, ksize, stride[0], stride[1], stride[2], pad, 0) == 1)
    # stride-1 3x3x3 with >= 64 channels per input: TMA halo + tcgen05 (conv_tc3.cu)
    use_tc3 = (use_tc and USE_TC3_CONV and w_tc.shape[-2] % 32 == 0
               and L.ltu_conv3d_tc3_supported(C0, C1, cout, ksize, stride[0], stride[1], stride[2], pad, int(up2),
                                              int(out_f32), n_aux) == 1)
    if use_tc3:
        use_halo = False
    if n_aux and not (use_tc or use_halo):
        raise RuntimeError("a fused auxiliary head needs the bf16 tensor-core path (tcgen05 or halo kernel)")
    aux = torch.empty(B, Ho, Wo, Do, n_aux, dtype=torch.float32, device=dev) if n_aux else None
    out = torch.empty(B, Ho, Wo, Do, cout, dtype=torch.float32 if out_f32 else x0.dtype, device=dev)
    if use_halo:
        tiles = L.ltu_conv3d_halo_tiles(Ho, Wo, Do, C0 + C1)
    elif use_tc3:
        tiles = L.ltu_conv3d_tc3_tiles(B, Hi, Wi, Di, cout, n_aux, int(up2))
    else:
        tiles = L.ltu_conv3d_tc_tiles(V, int(up2)) if use_tc else L.ltu_conv3d_tiles(V, cout)
    partials = torch.empty(B, tiles, cout, 2, dtype=torch.float32, device=dev) if want_stats else None
    cin = C0 + C1
    nbytes = (x0.numel() + (0 if x1 is None else x1.numel())) * x0.element_size() + out.numel() * out.element_size()
    # algorithmic flops: the un-folded count 2*k^3*Cin*Cout*B*V (the folded up2 path executes 8/27 of it)
    flops = 2 * ksize ** 3 * cin * cout * B * V
    folded = up2 and (use_tc or use_tc3)
    prof = ("conv3d_halo" if use_halo else ("conv3d_tc3" if use_tc3 else ("conv3d_tc" if use_tc else "conv3d")), nbytes,
            flops, flops * 8 // 27 if folded else flops)
    with _Guard(dev, prof) as st:
        if use_halo:
            check(L.ltu_conv3d_halo(_p(x0), C0, _p(x1), C1, B, Hi, Wi, Di, ksize, _p(w_tc), w_tc.shape[1], _p(bias),
                                    cout, _p(out), int(out_f32), _p(partials), n_aux, _p(aux), st), "ltu_conv3d_halo")
        elif use_tc3:
            check(L.ltu_conv3d_tc3(_p(x0), C0, _p(x1), C1, B, Hi, Wi, Di, int(up2), _p(w_tc), w_tc.shape[-2], w_tc.shape[-1],
                                   _p(bias), cout, _p(out), _p(partials), n_aux, _p(aux), st), "ltu_conv3d_tc3")
        elif use_tc:
            check(L.ltu_conv3d_tc(_p(x0), C0, _p(x1), C1, B, Hi, Wi, Di, int(up2), ksize, stride[0], stride[1],
                                  stride[2], pad, _p(w_tc), _p(bias), cout, _p(out), int(out_f32), Ho, Wo, Do,
                                  _p(partials), n_aux, _p(aux), st), "ltu_conv3d_tc")
        else:
            check(L.ltu_conv3d(_p(x0), C0, _p(x1), C1, B, Hi, Wi, Di, int(up2), ksize, stride[0], stride[1],
                               stride[2], pad, _p(w_packed), _p(bias), cout, _p(out), int(out_f32), Ho, Wo, Do,
                               _p(partials), _dt(x0), st), "ltu_conv3d")
    if n_aux:
        return out, partials, tiles, aux
    return out, partials, tiles


def instnorm_finalize(partials: torch.Tensor, voxels: int, eps: float = 1e-5) -> torch.Tensor:
    """partials [B,tiles,C,2] -> stats fp32 [B,C,2] = (mean, rstd)."""
    dev = _chk(partials)
    B, tiles, C, _ = partials.shape
    stats = torch.empty(B, C, 2, dtype=torch.float32, device=dev)
    with _Guard(dev, ("instnorm_finalize", partials.numel() * 4, 0)) as st:
        check(_native.lib().ltu_instnorm_finalize(_p(partials), _p(stats), B, tiles, C, voxels, eps, st),
              "ltu_instnorm_finalize")
    return stats


def chan_stats(x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """InstanceNorm statistics of an existing channels-last tensor [B,...,C]."""
    dev = _chk(x)
    B, C = x.shape[0], x.shape[-1]
    V = x.numel() // (B * C)
    chunks = max(1, min(256, V // 512))
    partials = torch.empty(B, chunks, C, 2, dtype=torch.float32, device=dev)
    with _Guard(dev) as st:
        check(_native.lib().ltu_chan_partials(_p(x), _p(partials), B, V, C, chunks, _dt(x), st), "ltu_chan_partials")
    return instnorm_finalize(partials, V, eps)


def instnorm_apply(x: torch.Tensor, stats: torch.Tensor, act: int = ACT_LRELU,
                   residual: Optional[torch.Tensor] = None, inplace: bool = True) -> torch.Tensor:
    """y = act((x - mean) * rstd) (+ residual)."""
    dev = _chk(x, stats, residual)
    B, C = x.shape[0], x.shape[-1]
    V = x.numel() // (B * C)
    y = x if inplace else torch.empty_like(x)
    nb = (2 + (residual is not None)) * x.numel() * x.element_size()
    with _Guard(dev, ("instnorm_apply", nb, 0)) as st:
        check(_native.lib().ltu_instnorm_apply(_p(x), _p(stats), _p(residual), _p(y), B, V, C, act, _dt(x), st),
              "ltu_instnorm_apply")
    return y


def zero_insert(y: torch.Tensor, size: Tuple[int, int, int], stride: Tuple[int, int, int]) -> torch.Tensor:
    """z [B,*size,C] with z[:, ::sh, ::sw, ::sd] = y and zeros elsewhere (input gradient of a strided convolution)."""
    dev = _chk(y)
    B, H, W, D, C = y.shape
    z = torch.empty(B, size[0], size[1], size[2], C, dtype=y.dtype, device=dev)
    with _Guard(dev) as st:
        check(_native.lib().ltu_zero_insert(_p(y), _p(z), B, H, W, D, C, size[0], size[1], size[2], stride[0], stride[1],
                                            stride[2], _dt(y), st), "ltu_zero_insert")
    return z


def sumpool2(x: torch.Tensor) -> torch.Tensor:
    """[B,2H,2W,2D,C] -> [B,H,W,D,C]: sums of 2x2x2 blocks (backward of the nearest x2 upsample)."""
    dev = _chk(x)
    B, H2, W2, D2, C = x.shape
    y = torch.empty(B, H2 // 2, W2 // 2, D2 // 2, C, dtype=x.dtype, device=dev)
    with _Guard(dev) as st:
        check(_native.lib().ltu_sumpool2(_p(x), _p(y), B, H2 // 2, W2 // 2, D2 // 2, C, _dt(x), st), "ltu_sumpool2")
    return y


def conv3d_wgrad(x: torch.Tensor, dy: torch.Tensor, ksize: int, stride: Tuple[int, int, int] = (1, 1, 1),
                 pad: int = 1, up2: bool = False) -> torch.Tensor:
    """Weight gradient of nn.Conv3d: x bf16 [B,Hi,Wi,Di,Cin], dy bf16 [B,Ho,Wo,Do,Cout] -> fp32 [k^3, Cout, Cin]
    (tap = (kh*3+kw)*3+kd).  As a parameter gradient: dw.view(k,k,k,Cout,Cin).permute(3,4,0,1,2).  `up2`: the
    convolution read the nearest-x2 upsampling of x."""
    dev = _chk(x, dy)
    if x.dtype != torch.bfloat16 or dy.dtype != torch.bfloat16:
        raise TypeError("conv3d_wgrad runs on the bf16 path (fp32 accumulation)")
    B, Hi, Wi, Di, Cin = x.shape
    _, Ho, Wo, Do, Cout = dy.shape
    L = _native.lib()
    nbytes = L.ltu_conv3d_wgrad_workspace(B, Ho, Wo, Do, Cin, Cout, ksize)
    ws = torch.empty(max(nbytes // 4, 1), dtype=torch.float32, device=dev)
    dw = torch.empty(ksize ** 3, Cout, Cin, dtype=torch.float32, device=dev)
    flops = 2 * ksize ** 3 * Cin * Cout * B * Ho * Wo * Do
    with _Guard(dev, ("conv3d_wgrad", (x.numel() + dy.numel()) * 2, flops)) as st:
        check(L.ltu_conv3d_wgrad(_p(x), _p(dy), _p(dw), _p(ws), nbytes, B, Hi, Wi, Di, Cin, Ho, Wo, Do, Cout, ksize,
                                 stride[0], stride[1], stride[2], pad, int(up2), st), "ltu_conv3d_wgrad")
    return dw


def instnorm_bwd(x_raw: torch.Tensor, stats: torch.Tensor, dy: torch.Tensor, act: int = ACT_LRELU) -> torch.Tensor:
    """Backward of act(InstanceNorm(x_raw)): x_raw [B,...,C] is the raw conv output, stats [B,C,2] the forward's
    (mean, rstd).  Returns dx; a residual added after the activation receives dy unchanged."""
    dev = _chk(x_raw, stats, dy)
    B, C = x_raw.shape[0], x_raw.shape[-1]
    V = x_raw.numel() // (B * C)
    L = _native.lib()
    nbytes = L.ltu_instnorm_bwd_workspace(B, V, C)
    ws = torch.empty(max(nbytes // 4, 1), dtype=torch.float32, device=dev)
    dx = torch.empty_like(x_raw)
    with _Guard(dev, ("instnorm_bwd", 5 * x_raw.numel() * x_raw.element_size(), 0)) as st:
        check(L.ltu_instnorm_bwd(_p(x_raw), _p(stats), _p(dy), _p(dx), _p(ws), nbytes, B, V, C, act, _dt(x_raw), st),
              "ltu_instnorm_bwd")
    return dx


# ------------------------------------------------------------------ U-Net plumbing
def s2d_input(x: torch.Tensor, dtype: torch.dtype, cpad: int = 4) -> torch.Tensor:
    """[B,1,H,W,D] fp32 -> [B,H/2,W/2,D,cpad] (windows_embedding, model/Unet_3Dblock.py:123-136);
    cpad=8 appends four zero channels (one 16-byte bf16 vector per voxel for the tensor-core stem)."""
    dev = _chk(x)
    if x.dtype != torch.float32:
        raise TypeError("model input must be float32")
    B, cin, H, W, D = x.shape
    if cin != 1:
        raise ValueError("windows_embedding requires dim_input == 1 (model/Unet_3Dblock.py:132)")
    y = torch.empty(B, H // 2, W // 2, D, cpad, dtype=dtype, device=dev)
    with _Guard(dev, ("s2d_input", x.numel() * 4, 0)) as st:
        check(_native.lib().ltu_s2d_input(_p(x), _p(y), B, H, W, D, cpad, _dt(y), st), "ltu_s2d_input")
    return y


def upsample_trilinear(x: torch.Tensor, fd: int) -> torch.Tensor:
    """Trilinear x(2,2,fd), align_corners=True (model/Unet_3Dblock.py:1341-1345)."""
    dev = _chk(x)
    B, H, W, D, C = x.shape
    y = torch.empty(B, 2 * H, 2 * W, fd * D, C, dtype=x.dtype, device=dev)
    with _Guard(dev, ("upsample", x.numel() * x.element_size() * (1 + 4 * fd), 0)) as st:
        check(_native.lib().ltu_upsample_trilinear(_p(x), _p(y), B, H, W, D, C, fd, _dt(x), st),
              "ltu_upsample_trilinear")
    return y


def upsample_trilinear_bwd(dy: torch.Tensor, fd: int) -> torch.Tensor:
    """Transpose of upsample_trilinear: dy [B,2H,2W,fd*D,C] -> dx [B,H,W,D,C]."""
    dev = _chk(dy)
    B, Ho, Wo, Do, C = dy.shape
    H, W, D = Ho // 2, Wo // 2, Do // fd
    dx = torch.empty(B, H, W, D, C, dtype=dy.dtype, device=dev)
    with _Guard(dev) as st:
        check(_native.lib().ltu_upsample_trilinear_bwd(_p(dy), _p(dx), B, H, W, D, C, fd, _dt(dy), st),
              "ltu_upsample_trilinear_bwd")
    return dx


def mask_softmax_bwd(logits: torch.Tensor, dmask: torch.Tensor) -> torch.Tensor:
    """logits fp32 [B,h,w,d,Cout], dmask fp32 [B,Cout,h,w,d] -> dlogits fp32 like logits."""
    dev = _chk(logits, dmask)
    B, h, w, d, C = logits.shape
    out = torch.empty_like(logits)
    with _Guard(dev) as st:
        check(_native.lib().ltu_mask_softmax_bwd(_p(logits), _p(dmask), _p(out), B, h * w * d, C, st), "ltu_mask_softmax_bwd")
    return out


def head_d2s_softmax_bwd(logits: torch.Tensor, dprobs: torch.Tensor, cout: int) -> torch.Tensor:
    """logits fp32 [B,H2,W2,D,4*cout], dprobs fp32 [B,cout,2*H2,2*W2,D] -> dlogits fp32 like logits."""
    dev = _chk(logits, dprobs)
    B, H2, W2, D, C4 = logits.shape
    assert C4 == 4 * cout
    out = torch.empty_like(logits)
    with _Guard(dev) as st:
        check(_native.lib().ltu_head_d2s_softmax_bwd(_p(logits), _p(dprobs), _p(out), B, H2, W2, D, cout, st),
              "ltu_head_d2s_softmax_bwd")
    return out


def mask_softmax(logits: torch.Tensor, want_mask: bool):
    """logits fp32 [B,h,w,d,Cout] -> (mask fp32 [B,Cout,h,w,d] | None, fg fp32 [B,h,w,d])."""
    dev = _chk(logits)
    B, h, w, d, C = logits.shape
    mask = torch.empty(B, C, h, w, d, dtype=torch.float32, device=dev) if want_mask else None
    fg = torch.empty(B, h, w, d, dtype=torch.float32, device=dev)
    with _Guard(dev, ("mask_softmax", logits.numel() * 4, 0)) as st:
        check(_native.lib().ltu_mask_softmax(_p(logits), _p(mask), _p(fg), B, h * w * d, C, st), "ltu_mask_softmax")
    return mask, fg


def gate_fused(a: torch.Tensor, stats_a: torch.Tensor, g: torch.Tensor, stats_g: torch.Tensor,
               psi_w: torch.Tensor, psi_b: torch.Tensor, skip: torch.Tensor) -> torch.Tensor:
    dev = _chk(a, stats_a, g, stats_g, psi_w, psi_b, skip)
    B, Ci = a.shape[0], a.shape[-1]
    V = a.numel() // (B * Ci)
    out = torch.empty_like(skip)
    with _Guard(dev, ("gate", 4 * a.numel() * a.element_size(), 0)) as st:
        check(_native.lib().ltu_gate_fused(_p(a), _p(stats_a), _p(g), _p(stats_g), _p(psi_w), _p(psi_b), _p(skip),
                                           _p(out), B, V, Ci, _dt(a), st), "ltu_gate_fused")
    return out


def gate_bwd(a: torch.Tensor, stats_a: torch.Tensor, g: torch.Tensor, stats_g: torch.Tensor, psi_w: torch.Tensor,
             psi_b: torch.Tensor, skip: torch.Tensor, dout: torch.Tensor):
    """Backward of gate_fused: returns (dskip_direct, dh, dpsi_w fp32 [Ci], dpsi_b fp32 [1]); dh is the gradient of
    BOTH normalised 1x1x1 conv outputs (pass it through instnorm_bwd(..., ACT_NONE) for each)."""
    dev = _chk(a, stats_a, g, stats_g, psi_w, psi_b, skip, dout)
    B, Ci = a.shape[0], a.shape[-1]
    V = a.numel() // (B * Ci)
    L = _native.lib()
    nbytes = L.ltu_gate_bwd_workspace(B, V, Ci)
    ws = torch.empty(max(nbytes // 4, 1), dtype=torch.float32, device=dev)
    dskip, dh = torch.empty_like(skip), torch.empty_like(a)
    dpw = torch.empty(Ci, dtype=torch.float32, device=dev)
    dpb = torch.empty(1, dtype=torch.float32, device=dev)
    with _Guard(dev) as st:
        check(L.ltu_gate_bwd(_p(a), _p(stats_a), _p(g), _p(stats_g), _p(psi_w), _p(psi_b), _p(skip), _p(dout), _p(dskip),
                             _p(dh), _p(dpw), _p(dpb), _p(ws), nbytes, B, V, Ci, _dt(a), st), "ltu_gate_bwd")
    return dskip, dh, dpw, dpb


def roi_bbox(fg: torch.Tensor, min_h: int, min_w: int, thr: float = 0.5) -> torch.Tensor:
    """fg fp32 [B,h,w,d] -> boxes fp32 [B,6] on device (model/Unet_3Dblock.py:821-873)."""
    dev = _chk(fg)
    B, h, w, d = fg.shape
    L = _native.lib()
    nbytes = L.ltu_roi_bbox_scratch(B, h, w)
    scratch = torch.empty(nbytes // 4, dtype=torch.int32, device=dev)
    box = torch.empty(B, 6, dtype=torch.float32, device=dev)
    with _Guard(dev) as st:
        check(L.ltu_roi_bbox(_p(fg), _p(box), _p(scratch), nbytes, B, h, w, d, min_h, min_w, thr, st), "ltu_roi_bbox")
    return box


def roi_resample(x: torch.Tensor, box: torch.Tensor, full_hw: Tuple[int, int], roi_h: int, roi_w: int,
                 eval_h: int, eval_w: int, direction: int) -> torch.Tensor:
    """Fisheye resample: direction 0 map->roi ([B,h,w,d,C] -> [B,eval_h,eval_w,d,C]), 1 back."""
    dev = _chk(x, box)
    B, _, _, d, C = x.shape
    h, w = full_hw
    oh, ow = (eval_h, eval_w) if direction == 0 else (h, w)
    y = torch.empty(B, oh, ow, d, C, dtype=x.dtype, device=dev)
    with _Guard(dev, ("roi_resample", 2 * x.numel() * x.element_size(), 0)) as st:
        check(_native.lib().ltu_roi_resample(_p(x), _p(box), _p(y), B, h, w, d, C, roi_h, roi_w, eval_h, eval_w,
                                             direction, _dt(x), st), "ltu_roi_resample")
    return y


def roi_resample_bwd(dy: torch.Tensor, box: torch.Tensor, full_hw: Tuple[int, int], roi_h: int, roi_w: int,
                     eval_h: int, eval_w: int, direction: int) -> torch.Tensor:
    """Transpose of roi_resample (same arguments): dy has the forward's output extent, the result its input extent."""
    dev = _chk(dy, box)
    B, _, _, d, C = dy.shape
    h, w = full_hw
    ih, iw = (h, w) if direction == 0 else (eval_h, eval_w)
    L = _native.lib()
    nbytes = L.ltu_roi_resample_bwd_workspace(B, h, w, eval_h, eval_w)
    ws = torch.empty(max(nbytes // 16, 1), 4, dtype=torch.float32, device=dev)
    dx = torch.empty(B, ih, iw, d, C, dtype=dy.dtype, device=dev)
    with _Guard(dev) as st:
        check(L.ltu_roi_resample_bwd(_p(dy), _p(box), _p(dx), _p(ws), nbytes, B, h, w, d, C, roi_h, roi_w, eval_h, eval_w,
                                     direction, _dt(dy), st), "ltu_roi_resample_bwd")
    return dx


def head_d2s_softmax(logits: torch.Tensor, cout: int, want_probs: bool, want_onehot: bool, want_labels: bool):
    """logits fp32 [B,H2,W2,D,4*cout] -> (probs, onehot, labels) in the reference layout."""
    dev = _chk(logits)
    B, H2, W2, D, C4 = logits.shape
    assert C4 == 4 * cout
    shp = (B, cout, 2 * H2, 2 * W2, D)
    probs = torch.empty(shp, dtype=torch.float32, device=dev) if want_probs else None
    onehot = torch.empty(shp, dtype=torch.float32, device=dev) if want_onehot else None
    labels = torch.empty((B, 2 * H2, 2 * W2, D), dtype=torch.uint8, device=dev) if want_labels else None
    with _Guard(dev, ("head_d2s_softmax", logits.numel() * 4, 0)) as st:
        check(_native.lib().ltu_head_d2s_softmax(_p(logits), _p(probs), _p(onehot), _p(labels), B, H2, W2, D, cout, st),
              "ltu_head_d2s_softmax")
    return probs, onehot, labels


def vote_accumulate(labels: torch.Tensor, starts: torch.Tensor, votes: torch.Tensor) -> None:
    """labels uint8 [n,rh,rw,rd], starts int32 [n,3], votes uint8 [C,H,W,D] (+= one-hot)."""
    dev = _chk(labels, starts, votes)
    n, rh, rw, rd = labels.shape
    C, H, W, D = votes.shape
    with _Guard(dev) as st:
        check(_native.lib().ltu_vote_accumulate(_p(labels), _p(starts), _p(votes), n, rh, rw, rd, C, H, W, D, st),
              "ltu_vote_accumulate")


def vote_argmax(votes: torch.Tensor) -> torch.Tensor:
    dev = _chk(votes)
    C, H, W, D = votes.shape
    out = torch.empty(H, W, D, dtype=torch.uint8, device=dev)
    with _Guard(dev) as st:
        check(_native.lib().ltu_vote_argmax(_p(votes), _p(out), C, H * W * D, st), "ltu_vote_argmax")
    return out


def vote_fractions(votes: torch.Tensor) -> torch.Tensor:
    """votes uint8 [C,H,W,D] -> fp32 [C,H,W,D] vote fractions (votes / number of covering windows)."""
    dev = _chk(votes)
    C, H, W, D = votes.shape
    out = torch.empty(C, H, W, D, dtype=torch.float32, device=dev)
    with _Guard(dev) as st:
        check(_native.lib().ltu_vote_fractions(_p(votes), _p(out), C, H * W * D, st), "ltu_vote_fractions")
    return out


DECIDE_THRESHOLD, DECIDE_ROUND = 0, 1


def vote_decide(votes: torch.Tensor, mode: int, thr: float = 0.5) -> torch.Tensor:
    """votes uint8 [C,H,W,D] -> uint8 one-hot [C,H,W,D]: `(fractions >= thr)` (inference_embed_attn.py:147) or
    `torch.round(fractions)` (inference_multi_classes.py:148) without materialising the fp32 fractions."""
    dev = _chk(votes)
    if votes.dtype != torch.uint8:
        raise TypeError("vote_decide: votes must be uint8")
    C = votes.shape[0]
    out = torch.empty_like(votes)
    with _Guard(dev) as st:
        check(_native.lib().ltu_vote_decide(_p(votes), _p(out), C, votes.numel() // C, mode, thr, st), "ltu_vote_decide")
    return out


def keep_largest_component_(onehot: torch.Tensor, applied_labels, connectivity: int = 3,
                            independent: bool = False) -> torch.Tensor:
    """In-place monai.transforms.KeepLargestConnectedComponent (0.7.0) on a uint8 one-hot volume [C,H,W,D]
    (inference_multi_classes.py:104,:150).  The labelling runs on the GPU (union-find), nothing goes to the host."""
    dev = _chk(onehot)
    if onehot.dtype != torch.uint8 or onehot.dim() != 4:
        raise TypeError("keep_largest_component_: onehot must be uint8 [C,H,W,D]")
    C, H, W, D = onehot.shape
    labels = [int(a) for a in applied_labels]
    if not labels or min(labels) < 0 or max(labels) >= C:
        raise ValueError("keep_largest_component_: applied_labels must be channel indices")
    L = _native.lib()
    nbytes = L.ltu_keep_largest_component_workspace(H * W * D)
    ws = torch.empty(nbytes // 8, dtype=torch.int64, device=dev)
    masks = [1 << a for a in labels] if independent else [sum(1 << a for a in set(labels))]
    with _Guard(dev) as st:
        for m in masks:
            check(L.ltu_keep_largest_component(_p(onehot), C, m, H, W, D, connectivity, _p(ws), nbytes, st),
                  "ltu_keep_largest_component")
    return onehot


def overlap_counts(pred_onehot: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """pred uint8 one-hot [C,H,W,D], target uint8 class indices [H,W,D] -> int64 [C+1,H,3] = (TP, predicted, target)
    per class and H-row; row C is the foreground pseudo-class (1 - pred[0]) vs (target != 0)."""
    dev = _chk(pred_onehot, target)
    if pred_onehot.dtype != torch.uint8 or target.dtype != torch.uint8:
        raise TypeError("overlap_counts: uint8 tensors expected")
    C, H, W, D = pred_onehot.shape
    if tuple(target.shape) != (H, W, D):
        raise ValueError("overlap_counts: target must be [H,W,D]")
    out = torch.empty(C + 1, H, 3, dtype=torch.int64, device=dev)
    with _Guard(dev) as st:
        check(_native.lib().ltu_overlap_counts(_p(pred_onehot), _p(target), C, H, W, D, _p(out), st), "ltu_overlap_counts")
    return out


def gather_windows(volume: torch.Tensor, starts: torch.Tensor, roi) -> torch.Tensor:
    """volume fp32 [H,W,D], starts int32 [n,3] -> windows fp32 [n,1,rh,rw,rd]."""
    dev = _chk(volume, starts)
    if volume.dtype != torch.float32 or starts.dtype != torch.int32:
        raise TypeError("gather_windows: volume must be float32 and starts int32")
    H, W, D = volume.shape
    n = starts.shape[0]
    rh, rw, rd = roi
    out = torch.empty(n, 1, rh, rw, rd, dtype=torch.float32, device=dev)
    with _Guard(dev) as st:
        check(_native.lib().ltu_gather_windows(_p(volume), _p(starts), _p(out), n, rh, rw, rd, H, W, D, st),
              "ltu_gather_windows")
    return out


EPI_BIAS, EPI_GELU, EPI_RES_LN = 0, 1, 2


def pack_linear_tc(weight: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin] -> bf16 [Cout16, Kpad64] K-major operand of ltu_linear_tc."""
    cout, cin = weight.shape
    out = torch.zeros((cout + 15) // 16 * 16, (cin + 63) // 64 * 64, dtype=torch.bfloat16, device=weight.device)
    out[:cout, :cin] = weight.detach().to(torch.bfloat16)
    return out


def linear_tc(x: torch.Tensor, w_packed: torch.Tensor, bias: torch.Tensor, cout: int, epi: int = EPI_BIAS,
              residual: Optional[torch.Tensor] = None, gamma: Optional[torch.Tensor] = None,
              beta: Optional[torch.Tensor] = None, eps: float = 1e-6) -> torch.Tensor:
    """y = epi(x @ W^T + b) on bf16 tokens [..., Cin] with the persistent tcgen05 kernel.  Layers wider
    than 256 outputs are computed as column slices of one output buffer."""
    dev = _chk(x, w_packed, bias, residual, gamma, beta)
    if x.dtype != torch.bfloat16:
        raise TypeError("linear_tc needs bf16 activations")
    cin = x.shape[-1]
    rows = x.numel() // cin
    y = torch.empty(*x.shape[:-1], cout, dtype=torch.bfloat16, device=dev)
    kpad = w_packed.shape[1]
    L = _native.lib()
    with _Guard(dev, ("linear_tc", (x.numel() + y.numel() + (0 if residual is None else residual.numel())) * 2,
                      2 * rows * cin * cout)) as st:
        for n0 in range(0, cout, 256):
            n = min(256, cout - n0)
            if epi == EPI_RES_LN and cout > 256:
                raise ValueError("the LayerNorm epilogue needs the whole row in one tile (Cout <= 256)")
            check(L.ltu_linear_tc(_p(x), cin, rows, c_void_p(w_packed.data_ptr() + n0 * kpad * 2),
                                  c_void_p(bias.data_ptr() + n0 * 4), n, c_void_p(y.data_ptr() + n0 * 2), cout, epi,
                                  _p(residual), _p(gamma), _p(beta), eps, st), "ltu_linear_tc")
    return y


def linear_fma(x: torch.Tensor, w_kn: torch.Tensor, bias: torch.Tensor, n: int) -> torch.Tensor:
    """nn.Linear on the fp32 path: y = x W^T + b as exact fp32 FMAs on CUDA cores -- the 1x1x1 case of ltu_conv3d (a
    shared-memory tiled GEMM), so the fp32 forward launches no library GEMM either.  x [..., K] contiguous fp32,
    w_kn = W^T as fp32 [1, K, N] (the [tap][Cin][Cout] packing of ltu_conv3d), bias fp32 [N]."""
    k = x.shape[-1]
    rows = x.numel() // k
    out, _, _ = conv3d(x.reshape(1, rows, 1, 1, k), w_kn, bias, n, 1, pad=0)
    return out.reshape(*x.shape[:-1], n)


def linear_fused_supported(k: int, n: int) -> bool:
    """Shapes ltu_linear_fused takes (the encoder layers with d_model 256, the K/V projection of d_model 128)."""
    return k % 64 == 0 and 64 <= k <= 1024 and n in (256, 512, 768)


def linear_fused(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, epi: int = EPI_BIAS,
                 res_hi: Optional[torch.Tensor] = None, res_lo: Optional[torch.Tensor] = None,
                 gamma: Optional[torch.Tensor] = None, beta: Optional[torch.Tensor] = None, eps: float = 1e-6,
                 want_lo: bool = True, softmax_cols: int = 0, x_cols: Optional[int] = None):
    """nn.Linear on bf16 tokens [..., K] with a fused epilogue, one persistent TMA + tcgen05 launch (csrc/linear_tma.cu):
    EPI_BIAS -> x W^T + b; EPI_GELU -> gelu(x W^T + b); EPI_RES_LN -> LayerNorm(x W^T + b + res_hi + res_lo) returned as the
    split pair (y_hi, y_lo) (y_lo None unless want_lo).  w: the nn.Linear weight [N, K] in bf16, bias / gamma / beta fp32.
    softmax_cols: output columns [0, softmax_cols) come out as softmax over each head's 32 columns / sqrt(32) (the Q third
    of a QKV projection, model/trans_block.py:50).  w of shape [B, N, K] with x [B, tokens, K]: one weight per sample
    (ctx_project: the readout folded into the output projection).  x_cols: the operand is columns [0, x_cols) of x (a
    column slice of a wider row, e.g. the Q third of a QKV tensor, read in place)."""
    dev = _chk(x, w, bias, res_hi, res_lo, gamma, beta)
    if x.dtype != torch.bfloat16 or w.dtype != torch.bfloat16:
        raise TypeError("linear_fused needs bf16 activations and bf16 weights")
    samples = 1
    if w.dim() == 3:
        samples = w.shape[0]
        if x.dim() != 3 or x.shape[0] != samples:
            raise ValueError(f"linear_fused: per-sample weights {tuple(w.shape)} need x [B, tokens, K], got {tuple(x.shape)}")
    n, k = w.shape[-2:]
    ldx = x.shape[-1]
    if (x_cols if x_cols is not None else ldx) != k or ldx < k:
        raise ValueError(f"linear_fused: x has {x.shape[-1]} columns (x_cols {x_cols}), the weight expects {k}")
    rows = x.numel() // ldx
    y = torch.empty(*x.shape[:-1], n, dtype=torch.bfloat16, device=dev)
    y_lo = torch.empty_like(y) if (epi == EPI_RES_LN and want_lo) else None
    nres = 0 if res_hi is None else (1 if res_lo is None else 2)
    nbytes = (rows * k + y.numel() * (2 if y_lo is not None else 1) + nres * rows * n) * 2
    with _Guard(dev, ("linear_fused", nbytes, 2 * rows * k * n)) as st:
        check(_native.lib().ltu_linear_fused_ex(_p(x), ldx, rows, k, _p(w), _p(bias), n, epi, _p(res_hi), _p(res_lo), _p(gamma),
                                                _p(beta), eps, _p(y), _p(y_lo), softmax_cols, samples, st), "ltu_linear_fused_ex")
    return (y, y_lo) if epi == EPI_RES_LN else y


def ctx_project(ctx: torch.Tensor, w_o: torch.Tensor) -> torch.Tensor:
    """ctx fp32 [B, heads, 32, 32] (kv_reduce) and the output projection's bf16 weight [C, C] -> bf16 [B, C, C],
    W_b[n, 32h + j] = sum_e ctx[b, h, j, e] Wo[n, 32h + e]: linear_fused(softmax(Q), W_b) is readout + projection
    (model/trans_block.py:65, :166)."""
    dev = _chk(ctx, w_o)
    B, heads = ctx.shape[:2]
    C = heads * 32
    if ctx.dtype != torch.float32 or w_o.dtype != torch.bfloat16 or tuple(w_o.shape) != (C, C) or tuple(ctx.shape[2:]) != (32, 32):
        raise TypeError("ctx_project needs ctx fp32 [B, heads, 32, 32] and the bf16 [C, C] weight with C = 32 heads")
    out = torch.empty(B, C, C, dtype=torch.bfloat16, device=dev)
    with _Guard(dev, ("ctx_project", B * C * C * 2, 2 * B * C * C * 32)) as st:
        check(_native.lib().ltu_ctx_project(_p(ctx), _p(w_o), _p(out), B, heads, st), "ltu_ctx_project")
    return out


def kv_project_reduce_supported(c: int, heads: int, n: int) -> bool:
    return bool(_native.lib().ltu_kv_project_reduce_supported(c, heads, n))


def kv_project_reduce(x: torch.Tensor, w_kv: torch.Tensor, b_kv: torch.Tensor, heads: int, w_o: Optional[torch.Tensor] = None):
    """ctx fp32 [B, heads, 32, 32] = kv_reduce(x Wk^T + bk, x Wv^T + bv) in one launch: the K/V projection's output tiles
    are reduced in the GEMM epilogue, K and V never reach memory (csrc/kv_project.cu; d_model 128, 4 heads, bf16).
    x [B, N, 128] bf16, w_kv [256, 128] bf16 (Wk rows, then Wv rows), b_kv fp32 [256].  With w_o (the output projection's
    bf16 weight [128, 128]) the merge kernel also writes ctx_project's per-sample weight and the result is (ctx, W_b)."""
    dev = _chk(x, w_kv, b_kv, w_o)
    B, N, C = x.shape
    if x.dtype != torch.bfloat16 or w_kv.dtype != torch.bfloat16 or tuple(w_kv.shape) != (2 * C, C):
        raise TypeError("kv_project_reduce needs bf16 tokens and the bf16 [2C, C] K/V weight")
    L = _native.lib()
    nbytes = L.ltu_kv_project_reduce_workspace(B, N)
    ws = torch.empty(max(nbytes // 4, 1), dtype=torch.float32, device=dev)
    ctx = torch.empty(B, heads, 32, 32, dtype=torch.float32, device=dev)
    # algorithmic bytes = SURVEY 8(d)'s figure for the op this launch implements (kv_reduce: K and V read once, 2*B*N*C*E);
    # what it actually moves is half of that (x once: K is never stored, V never computed per token)
    with _Guard(dev, ("kv_project_reduce", 2 * B * N * C * 2, 2 * B * N * C * (2 * C + 32))) as st:
        wb = None
        if w_o is not None:
            if w_o.dtype != torch.bfloat16 or tuple(w_o.shape) != (C, C):
                raise TypeError("kv_project_reduce(w_o=...) needs the bf16 [C, C] weight")
            wb = torch.empty(B, C, C, dtype=torch.bfloat16, device=dev)
        check(L.ltu_kv_project_reduce(_p(x), _p(w_kv), _p(b_kv), _p(ctx), _p(ws), nbytes, B, N, _p(w_o), _p(wb), st),
              "ltu_kv_project_reduce")
    return ctx if w_o is None else (ctx, wb)


def ffn_fused_supported(c: int) -> bool:
    return bool(_native.lib().ltu_ffn_fused_supported(c))


def ffn_fused(x: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor,
              gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-6, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """LayerNorm(x + W2 gelu(W1 x + b1) + b2) on bf16 tokens [..., C] in one persistent tcgen05 kernel
    (model/trans_block.py:207-210).  w1 [2C, C] / w2 [C, 2C] bf16, the rest fp32."""
    dev = _chk(x, w1, b1, w2, b2, gamma, beta, out)
    c = x.shape[-1]
    if x.dtype != torch.bfloat16 or w1.dtype != torch.bfloat16 or w2.dtype != torch.bfloat16:
        raise TypeError("ffn_fused needs bf16 activations and bf16 weights")
    if tuple(w1.shape) != (2 * c, c) or tuple(w2.shape) != (c, 2 * c):
        raise ValueError(f"ffn_fused: weight shapes {tuple(w1.shape)} / {tuple(w2.shape)} do not match d_model {c}")
    rows = x.numel() // c
    y = torch.empty_like(x) if out is None else out
    with _Guard(dev, ("ffn_fused", 2 * x.numel() * 2, 2 * rows * c * 2 * c * 2)) as st:
        check(_native.lib().ltu_ffn_fused(_p(x), rows, c, _p(w1), _p(b1), _p(w2), _p(b2), _p(gamma), _p(beta), eps,
                                          _p(y), st), "ltu_ffn_fused")
    return y


def attn_out_fused_supported(c: int, heads: int) -> bool:
    return bool(_native.lib().ltu_attn_out_fused_supported(c, heads))


def ctx_pack(ctx: torch.Tensor) -> torch.Tensor:
    """fp32 ctx [B, heads, 32, 32] (kv_reduce) -> bf16 [B*heads*32, 64] tensor-core operand of attn_out_fused."""
    dev = _chk(ctx)
    B, heads = ctx.shape[0], ctx.shape[1]
    out = torch.empty(B * heads * 32, 64, dtype=torch.bfloat16, device=dev)
    with _Guard(dev, ("ctx_pack", ctx.numel() * 4, 0)) as st:
        check(_native.lib().ltu_ctx_pack_bf16(_p(ctx), _p(out), B, heads, st), "ltu_ctx_pack_bf16")
    return out


def attn_out_fused(x: torch.Tensor, wq: torch.Tensor, bq: torch.Tensor, ctx16: torch.Tensor, wo: torch.Tensor,
                   bo: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, heads: int, eps: float = 1e-6) -> torch.Tensor:
    """LayerNorm(x + readout(x Wq^T + bq, ctx) Wo^T + bo) on bf16 tokens [B, N, C] in one persistent tcgen05 kernel
    (model/trans_block.py:50,:65,:155-166,:205-206)."""
    dev = _chk(x, wq, bq, ctx16, wo, bo, gamma, beta)
    B, N, C = x.shape
    if x.dtype != torch.bfloat16 or wq.dtype != torch.bfloat16 or wo.dtype != torch.bfloat16 or ctx16.dtype != torch.bfloat16:
        raise TypeError("attn_out_fused needs bf16 activations, weights and packed ctx")
    if tuple(wq.shape) != (C, C) or tuple(wo.shape) != (C, C) or tuple(ctx16.shape) != (B * heads * 32, 64):
        raise ValueError("attn_out_fused: operand shapes do not match")
    y = torch.empty_like(x)
    with _Guard(dev, ("attn_out_fused", 2 * x.numel() * 2, 2 * B * N * C * (2 * C + 32))) as st:
        check(_native.lib().ltu_attn_out_fused(_p(x), B, N, C, heads, _p(wq), _p(bq), _p(ctx16), _p(wo), _p(bo),
                                               _p(gamma), _p(beta), eps, _p(y), st), "ltu_attn_out_fused")
    return y


def attn_out_fused_w(x: torch.Tensor, wq: torch.Tensor, bq: torch.Tensor, w_b: torch.Tensor, bo: torch.Tensor,
                     gamma: torch.Tensor, beta: torch.Tensor, heads: int, eps: float = 1e-6) -> torch.Tensor:
    """attn_out_fused with the readout folded into the output projection: w_b bf16 [B, C, C] = blockdiag(ctx_b) Wo^T
    (ctx_project / the merge kernel of kv_project_reduce, kv_reduce).  LayerNorm(x + softmax(x Wq^T + bq)/sqrt(32) W_b^T + bo)."""
    dev = _chk(x, wq, bq, w_b, bo, gamma, beta)
    B, N, C = x.shape
    if x.dtype != torch.bfloat16 or wq.dtype != torch.bfloat16 or w_b.dtype != torch.bfloat16:
        raise TypeError("attn_out_fused_w needs bf16 activations and weights")
    if tuple(wq.shape) != (C, C) or tuple(w_b.shape) != (B, C, C):
        raise ValueError("attn_out_fused_w: operand shapes do not match")
    y = torch.empty_like(x)
    with _Guard(dev, ("attn_out_fused", 2 * x.numel() * 2, 2 * B * N * C * 2 * C)) as st:
        check(_native.lib().ltu_attn_out_fused_w(_p(x), B, N, C, heads, _p(wq), _p(bq), _p(w_b), _p(bo), _p(gamma), _p(beta),
                                                 eps, _p(y), st), "ltu_attn_out_fused_w")
    return y
