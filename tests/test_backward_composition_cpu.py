"""CPU: the HOST LOGIC of lintransunet_b200/backward.py -- which gradient goes where, in which layout, under which
parameter name -- checked end to end without a GPU: every kernel wrapper in lintransunet_b200.ops is replaced by a torch
stand-in (forward = the op's definition in torch ops, backward = torch autograd of that forward, fp64), and
``model_loss_and_gradients`` must then reproduce the loss terms and the parameter gradients of fp64 autograd through the
oracle model + oracle loss.  The kernels themselves are verified one by one on the GPU (tests/test_*_bwd_gpu.py); this test
covers the composition, including the decoder loop that has not run on a GPU yet."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import ltu_oracle as O
from oracle import train_step as T

DT = torch.float64


def nc(t):   # channels-last [B,H,W,D,C] -> [B,C,H,W,D]
    return t.permute(0, 4, 1, 2, 3)


def cl(t):
    return t.permute(0, 2, 3, 4, 1).contiguous()


def vjp(fn, inputs, dy):
    """Gradients of fn(*inputs) w.r.t. the inputs, contracted with dy."""
    xs = [t.detach().clone().requires_grad_(True) for t in inputs]
    with torch.enable_grad():
        y = fn(*xs)
        return torch.autograd.grad(y, xs, dy.to(y.dtype))


class StandIns:
    """Torch definitions of the ops backward.py calls (same signatures as lintransunet_b200.ops)."""
    ACT_NONE, ACT_LRELU = 0, 1

    # ---- convolution + InstanceNorm
    @staticmethod
    def _conv(x, w_packed, bias, ksize, stride, pad, up2):
        taps, cin, cout = w_packed.shape
        w = w_packed.reshape(ksize, ksize, ksize, cin, cout).permute(4, 3, 0, 1, 2).to(x.dtype)
        xn = nc(x)
        if up2:
            xn = F.interpolate(xn, scale_factor=2, mode="nearest")
        return cl(F.conv3d(xn, w, None if bias is None else bias.to(x.dtype), stride=stride, padding=pad))

    @staticmethod
    def conv3d(x0, w_packed, bias, cout, ksize, stride=(1, 1, 1), pad=1, x1=None, up2=False, out_f32=False,
               want_stats=False, w_tc=None, w_tc_fold=None, n_aux=0):
        x = x0 if x1 is None else torch.cat([x0, x1], -1)
        out = StandIns._conv(x, w_packed, bias, ksize, stride, pad, up2)
        return out, (out if want_stats else None), 0          # "partials" = the raw output itself

    @staticmethod
    def instnorm_finalize(partials, voxels, eps=1e-5):
        mean = partials.mean((1, 2, 3))
        var = partials.var((1, 2, 3), unbiased=False)
        return torch.stack([mean, torch.rsqrt(var + eps)], -1)

    @staticmethod
    def _norm_act(x, act):
        # written out: the CPU backward of F.instance_norm is wrong for permuted (non-contiguous) tensors in this torch build
        y = (x - x.mean((1, 2, 3), keepdim=True)) * torch.rsqrt(x.var((1, 2, 3), unbiased=False, keepdim=True) + 1e-5)
        return F.leaky_relu(y, 0.01) if act == StandIns.ACT_LRELU else y

    @staticmethod
    def instnorm_apply(x, stats, act=1, residual=None, inplace=True):
        y = StandIns._norm_act(x, act)
        return y if residual is None else y + residual

    @staticmethod
    def instnorm_bwd(x_raw, stats, dy, act=1):
        return vjp(lambda x: StandIns._norm_act(x, act), [x_raw], dy)[0]

    @staticmethod
    def conv3d_wgrad(x, dy, ksize, stride=(1, 1, 1), pad=1, up2=False):
        cin, cout = x.shape[-1], dy.shape[-1]
        w0 = torch.zeros(ksize ** 3, cin, cout, dtype=x.dtype)
        dw = vjp(lambda w: StandIns._conv(x, w, None, ksize, stride, pad, up2), [w0], dy)[0]     # [taps, cin, cout]
        return dw.permute(0, 2, 1).contiguous()                                                   # [taps, cout, cin]

    @staticmethod
    def zero_insert(y, size, stride):
        B, H, W, D, C = y.shape
        z = torch.zeros(B, size[0], size[1], size[2], C, dtype=y.dtype)
        z[:, ::stride[0], ::stride[1], ::stride[2]][:, :H, :W, :D] = y
        return z

    @staticmethod
    def sumpool2(x):
        B, H2, W2, D2, C = x.shape
        return x.reshape(B, H2 // 2, 2, W2 // 2, 2, D2 // 2, 2, C).sum((2, 4, 6))

    @staticmethod
    def s2d_input(x, dtype, cpad=4):
        a = cl(O.space_to_depth(x.to(DT), 2))
        return torch.cat([a, torch.zeros(*a.shape[:-1], cpad - 4, dtype=DT)], -1) if cpad > 4 else a

    # ---- transformer
    @staticmethod
    def _heads(t, h):
        B, N, C = t.shape
        return t.reshape(B, N, h, 32).transpose(1, 2)

    @staticmethod
    def kv_reduce(k, v, heads):
        return torch.softmax(StandIns._heads(k, heads), -2).transpose(-1, -2) @ StandIns._heads(v, heads)

    @staticmethod
    def q_readout(q, ctx, heads):
        B, N, C = q.shape
        return ((torch.softmax(StandIns._heads(q, heads), -1) / math.sqrt(32)) @ ctx).transpose(1, 2).reshape(B, N, C)

    @staticmethod
    def linear_attention_bwd(q, k, v, ctx, dout, heads):
        B, N, C = q.shape

        def fwd(qkv):
            qq, kk, vv = (StandIns._heads(qkv[..., i * C:(i + 1) * C], heads) for i in range(3))
            return O.efficient_attention(qq, kk, vv).transpose(1, 2).reshape(B, N, C)
        return vjp(fwd, [torch.cat([q, k, v], -1)], dout)[0]

    @staticmethod
    def add_layernorm(x, res, gamma, beta, eps=1e-6):
        return F.layer_norm(x + res, (x.shape[-1],), gamma.to(x.dtype), beta.to(x.dtype), eps)

    @staticmethod
    def add_layernorm_bwd(x, res, dy, gamma, eps=1e-6):
        C = x.shape[-1]
        dx, dg, db = vjp(lambda a, g, b: F.layer_norm(a + res, (C,), g, b, eps), [x, gamma.to(x.dtype), torch.zeros(C, dtype=x.dtype)], dy)
        return dx, dg, db

    gelu = staticmethod(F.gelu)

    @staticmethod
    def gelu_bwd(x, dy):
        return vjp(F.gelu, [x], dy)[0]

    @staticmethod
    def _posenc(x, w27c, bias):
        C = x.shape[-1]
        w = w27c.reshape(3, 3, 3, C).permute(3, 0, 1, 2).unsqueeze(1).to(x.dtype)
        return x + cl(F.conv3d(nc(x), w, bias.to(x.dtype), padding=1, groups=C))

    posenc_dwconv3 = staticmethod(lambda x, w, b: StandIns._posenc(x, w, b))

    @staticmethod
    def posenc_dwconv3_bwd(x, dy, w27c):
        C = x.shape[-1]
        return vjp(StandIns._posenc, [x, w27c.to(x.dtype), torch.zeros(C, dtype=x.dtype)], dy)

    # ---- decoder glue
    @staticmethod
    def _up(x, fd):
        return cl(F.interpolate(nc(x), scale_factor=(2, 2, fd), mode="trilinear", align_corners=True))

    upsample_trilinear = staticmethod(lambda x, fd: StandIns._up(x, fd))

    @staticmethod
    def upsample_trilinear_bwd(dy, fd):
        B, Ho, Wo, Do, C = dy.shape
        x0 = torch.zeros(B, Ho // 2, Wo // 2, Do // fd, C, dtype=dy.dtype)
        return vjp(lambda x: StandIns._up(x, fd), [x0], dy)[0]

    @staticmethod
    def mask_softmax(logits, want_mask):
        p = torch.softmax(logits, -1)
        return (nc(p).contiguous() if want_mask else None), 1 - p[..., 0]

    @staticmethod
    def mask_softmax_bwd(logits, dmask):
        return vjp(lambda l: nc(torch.softmax(l, -1)), [logits], dmask)[0]

    @staticmethod
    def _head(logits):
        return torch.softmax(O.depth_to_space(nc(logits), 2), 1)

    @staticmethod
    def head_d2s_softmax(logits, cout, want_probs, want_onehot, want_labels):
        return StandIns._head(logits), None, None

    @staticmethod
    def head_d2s_softmax_bwd(logits, dprobs, cout):
        return vjp(StandIns._head, [logits], dprobs)[0]

    @staticmethod
    def _gate(a, g, psi_w, psi_b, skip):
        h = F.relu(StandIns._norm_act(a, 0) + StandIns._norm_act(g, 0))
        z = (h * psi_w.to(h.dtype)).sum(-1, keepdim=True) + psi_b.to(h.dtype)
        return skip * torch.sigmoid(z), h, z

    @staticmethod
    def gate_fused(a, sa, g, sg, psi_w, psi_b, skip):
        return StandIns._gate(a, g, psi_w, psi_b, skip)[0]

    @staticmethod
    def gate_bwd(a, sa, g, sg, psi_w, psi_b, skip, dout):
        _, h, z = StandIns._gate(a, g, psi_w, psi_b, skip)
        s = torch.sigmoid(z)
        dz = (dout * skip).sum(-1, keepdim=True) * s * (1 - s)
        dh = dz * psi_w.to(h.dtype) * (h > 0)
        return dout * s, dh, (dz * h).sum((0, 1, 2, 3)), dz.sum().reshape(1)

    @staticmethod
    def roi_bbox(fg, min_h, min_w, thr=0.5):
        return O.roi_boxes(fg.unsqueeze(1).float(), min_h, min_w, thr)

    @staticmethod
    def _resample(x, box, full_hw, roi_h, roi_w, eval_h, eval_w, direction):
        h, w = full_hw
        x0, y0, x1, y1 = box[:, 0:1], box[:, 1:2], box[:, 3:4], box[:, 4:5]
        f = O.fisheye_forward_coords if direction == 0 else O.fisheye_back_coords
        return cl(O.separable_resample(nc(x), f(x0, x1, h - 1, roi_h, eval_h), f(y0, y1, w - 1, roi_w, eval_w)))

    roi_resample = staticmethod(lambda x, box, full_hw, rh, rw, eh, ew, direction: StandIns._resample(x, box, full_hw, rh, rw, eh, ew, direction))

    @staticmethod
    def roi_resample_bwd(dy, box, full_hw, roi_h, roi_w, eval_h, eval_w, direction):
        B, _, _, d, C = dy.shape
        ih, iw = full_hw if direction == 0 else (eval_h, eval_w)
        x0 = torch.zeros(B, ih, iw, d, C, dtype=dy.dtype)
        return vjp(lambda x: StandIns._resample(x, box, full_hw, roi_h, roi_w, eval_h, eval_w, direction), [x0], dy)[0]


def _case():
    cfg = O.UnetConfig(dim_output=2)
    sd = O.make_state_dict(cfg, seed=0)
    x = O.make_input((1, 1, 64, 64, 16), seed=1, blob=True)
    masks = torch.zeros(1, 1, 64, 64, 16, dtype=torch.long)
    masks[:, :, 20:44, 16:40, 4:12] = 1
    return cfg, sd, x, masks


@pytest.fixture(scope="module")
def oracle_step():
    """fp64 autograd through the oracle model and the oracle loss, computed once for the module."""
    cfg, sd, x, masks = _case()
    sdd = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    out = O.mask_trans_unet_forward(x.double(), sdd, cfg)
    total, terms = T.train_loss(out["probs"], out["mask_list"], masks)
    total.backward()
    return dict(total=float(total.detach()), terms=[[float(v.detach()) for v in row] for row in terms],
                grads={k: v.grad for k, v in sdd.items()})


def _loss_sums_torch(p, labels):
    """Definition of ltu_loss_sums in torch ops (differentiable in p)."""
    n, c = p.shape[0], p.shape[1]
    pf = p.reshape(n, c, -1)
    lab = labels.reshape(n, 1, -1).long()
    onehot = (lab == torch.arange(c).view(1, c, 1)).to(p.dtype)
    s = -(1 - pf) * onehot * torch.log(torch.clamp(pf, min=1e-6))
    return torch.stack([pf.sum(-1), onehot.sum(-1), (pf * onehot).sum(-1), s.sum(-1)], -1)


def _install_loss_standins(monkeypatch):
    from lintransunet_b200 import losses, ops
    monkeypatch.setattr(ops, "loss_sums", _loss_sums_torch)
    monkeypatch.setattr(ops, "loss_sums_bwd", lambda p, labels, g: vjp(lambda q: _loss_sums_torch(q, labels), [p], g)[0])
    monkeypatch.setattr(ops, "label_pool", lambda lab, k: F.max_pool3d(lab.float().unsqueeze(1), kernel_size=k, stride=k)
                        .squeeze(1).to(torch.uint8))
    # keep fp64 probabilities in fp64 (the product casts to fp32 for the kernel)
    monkeypatch.setattr(losses, "level_sums", lambda predict, lab: (losses._LossSums.apply(predict, lab),
                                                                    predict.numel() // (predict.shape[0] * predict.shape[1])))


@pytest.fixture()
def standins(monkeypatch):
    from lintransunet_b200 import backward, ops
    for name in dir(StandIns):
        if not name.startswith("_") and name not in ("ACT_NONE", "ACT_LRELU"):
            monkeypatch.setattr(ops, name, getattr(StandIns, name))
    monkeypatch.setattr(backward, "_ACT", DT)
    _install_loss_standins(monkeypatch)
    return backward


def test_whole_model_gradient_composition_matches_oracle_autograd(standins, oracle_step):
    from lintransunet_b200 import MaskTransUnet
    cfg, sd, x, masks = _case()
    model = MaskTransUnet(list(cfg.num_layers), list(cfg.roi_size_list), list(cfg.is_roi_list), 1, 2, dropout=0.0)
    model.load_state_dict(sd)
    total, terms, grads = standins.model_loss_and_gradients(model, x, masks)
    assert abs(float(total) - oracle_step["total"]) <= 1e-6 * abs(oracle_step["total"])
    for row, row_ref in zip(terms, oracle_step["terms"]):
        for a, b in zip(row, row_ref):
            assert abs(float(a.detach()) - b) <= 1e-6 * max(1.0, abs(b))
    ref_grads = oracle_step["grads"]
    live = sorted(k for k, v in ref_grads.items() if v is not None)
    assert sorted(grads) == live                               # same 600 parameters, same names
    gmax = max(float(ref_grads[k].norm()) for k in live)
    worst, worst_name = 0.0, ""
    for k in live:
        ref = ref_grads[k]
        assert grads[k].shape == ref.shape, k
        err = float((grads[k].double() - ref).norm())
        if k.endswith(".bias"):
            # bias gradients are row sums taken in fp32 by backward.py; the mathematically zero ones (in front of an
            # InstanceNorm, K projection) are 1e-8 here and 1e-16 in the reference: absolute bound
            assert err <= 1e-6 * float(ref.norm()) + 1e-7 * gmax, (k, err)
            continue
        if float(ref.norm()) == 0.0:                           # e.g. the gates whose skip is replaced by a degenerate ROI bridge
            assert err <= 1e-7 * gmax, (k, err)
            continue
        e = err / max(float(ref.norm()), 1e-9 * gmax)          # layer 0 of the ROI bridges: 1e-18 gradients (constant input channel)
        if e > worst:
            worst, worst_name = e, k
    print(f"\n[backward composition, fp64 stand-ins] worst relative weight-gradient error {worst:.2e} ({worst_name})")
    assert worst <= 1e-6                                       # parameter gradients are returned in fp32


def test_autograd_wiring_fills_parameter_gradients(standins, oracle_step):
    """loss.backward() through lintransunet_b200.unet._NativeTrainFunction (what MaskTransUnet.forward returns in training
    mode with model.native_backward) gives every live parameter the oracle-autograd gradient and leaves the dead ones None."""
    from lintransunet_b200 import MaskTransUnet, losses
    from lintransunet_b200.unet import _NativeTrainFunction
    cfg, sd, x, masks = _case()
    model = MaskTransUnet(list(cfg.num_layers), list(cfg.roi_size_list), list(cfg.is_roi_list), 1, 2, dropout=0.0)
    model.load_state_dict(sd)
    model.train()
    out = _NativeTrainFunction.apply(model, x, *[p for _, p in model.named_parameters()])
    probs, mask_list = out[0], list(out[1:])
    assert probs.requires_grad and len(mask_list) == 4
    total, _ = losses.deep_supervision_loss(probs, mask_list, masks)
    total.backward()
    ref_grads = oracle_step["grads"]
    gmax = max(float(v.norm()) for v in ref_grads.values() if v is not None)
    for name, p in model.named_parameters():
        r = ref_grads[name]
        if r is None:
            assert p.grad is None, name
            continue
        assert p.grad is not None and p.grad.dtype == p.dtype and p.grad.shape == p.shape, name
        assert float((p.grad.double() - r).norm()) <= 1e-5 * float(r.norm()) + 1e-7 * gmax, name


def test_train_step_updates_parameters_like_an_oracle_step(standins, oracle_step, monkeypatch):
    """lintransunet_b200.train.train_step (forward, loss, backward through the autograd wrapper, SGD step) moves every
    parameter exactly as an SGD step on the oracle-autograd gradients."""
    from lintransunet_b200 import MaskTransUnet, train
    cfg, sd, x, masks = _case()
    model = MaskTransUnet(list(cfg.num_layers), list(cfg.roi_size_list), list(cfg.is_roi_list), 1, 2, dropout=0.0)
    model.load_state_dict(sd)
    monkeypatch.setattr(MaskTransUnet, "forward", lambda self, x: self._forward_train(x))      # skip the CUDA-only guard
    lr = 0.1
    opt = torch.optim.SGD(model.parameters(), lr=lr)
    total, terms = train.train_step(model, opt, x, masks)
    assert abs(total - oracle_step["total"]) <= 1e-6 * abs(oracle_step["total"]) and len(terms) == 5
    for name, p in model.named_parameters():
        g = oracle_step["grads"][name]
        want = sd[name].double() - (lr * g if g is not None else 0)
        assert torch.allclose(p.detach().double(), want, rtol=0, atol=1e-6 * float(want.abs().max()) + 1e-9), name
        assert p.grad is None or float(p.grad.abs().max()) == 0.0                 # zero_grad after the step
