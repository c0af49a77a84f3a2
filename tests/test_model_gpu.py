"""GPU parity of the whole drop-in MaskTransUnet forward (a5-a16) against (1) the vectors
produced by the unmodified reference (tests/golden, tools/make_golden.py) and (2) the CPU oracle
run on the same seeded weights/inputs.

Tolerances, as max|a-b|/max|ref| on the `decode.final_block` logits tap:
  fp32 path : <= 1e-3 (north star), argmax one-hot mismatch <= 1e-4 of voxels, ROI boxes bit-equal
  bf16 path : the reference's own bf16-autocast forward is 4.5e-2..7.6e-2 away from its fp32
              forward with 0.8..3.1 % argmax flips on random-init weights, so the
              north-star 2e-2 / 1e-4 cannot be a property of any bf16 pipeline on these near-tied
              logits.  Asserted here: <= 8e-2 on logits (measured 4e-2..6e-2), <= 3 % argmax flips, and <= 2e-2 /
              <= 1e-2 on the margin-filtered voxels (|top1-top2| logit gap > 0.25) -- measured
              values are printed and recorded in DESIGN.md.
"""
import numpy as np
import pytest
import torch

from oracle import ltu_oracle as O
from tests.helpers import load_golden, rel_err, sub

pytestmark = pytest.mark.gpu

CASES = ["c2_64x64x16", "c3_64x96x32_b2", "c2_384x384x16_wellformed"]


def from_cl(t):
    return t.permute(0, 4, 1, 2, 3).contiguous()


def build(g, precision):
    from lintransunet_b200 import MaskTransUnet
    cfg = O.UnetConfig(dim_output=int(g["dim_output"]))
    sd = O.make_state_dict(cfg, seed=int(g["seed_w"]))
    m = MaskTransUnet(list(cfg.num_layers), list(cfg.roi_size_list), list(cfg.is_roi_list), 1, cfg.dim_output,
                      dropout=0.0)
    m.load_state_dict(sd, strict=True)
    m.cuda()
    m.precision = precision
    x = O.make_input(tuple(int(s) for s in g["shape"]), seed=int(g["seed_x"]), blob=bool(g["blob"]))
    return m, cfg, sd, x


@pytest.mark.parametrize("name", CASES)
def test_fp32_forward_matches_reference_golden(name):
    g = load_golden(f"model_{name}.npz")
    m, cfg, sd, x = build(g, "fp32")
    m.record = {}
    m.train()
    with torch.no_grad():
        probs, mask_list = m(x.cuda())
    rec = m.record
    m.record = None
    m.eval()
    onehot = m(x.cuda())
    report = {}
    for i in (1, 2, 3):
        assert np.array_equal(rec[f"box{i}"].cpu().numpy(), g[f"box{i}"]), f"ROI box {i}"
    report["bottle"] = rel_err(sub(from_cl(rec["bottle"])), g["bottle"])
    for i in range(4):
        report[f"skip{i}"] = rel_err(sub(from_cl(rec[f"skip{i}"])), g[f"skip{i}"])
        report[f"up{i}"] = rel_err(sub(from_cl(rec[f"up{i}"])), g[f"up{i}"])
        report[f"mask{i}"] = rel_err(sub(mask_list[i]), g[f"mask{i}"])
    for i in (1, 2, 3, 4):
        report[f"bridge{i}"] = rel_err(sub(from_cl(rec[f"bridge{i}"])), g[f"bridge{i}"])
    report["logits"] = rel_err(sub(from_cl(rec["logits"])), g["logits"])
    report["probs"] = rel_err(sub(probs), g["probs"])
    report["onehot_mismatch"] = float(np.mean(sub(onehot) != g["onehot"]))
    print(f"\n[fp32 {name}] " + " ".join(f"{k}={v:.2e}" for k, v in report.items()))
    assert onehot.shape == tuple(int(s) for s in (g["shape"][0], g["dim_output"], *g["shape"][2:]))
    assert float(onehot.sum(1).min()) == 1.0 and float(onehot.sum(1).max()) == 1.0
    for k, v in report.items():
        assert v < (1e-4 if k == "onehot_mismatch" else 1e-3), (k, v)


@pytest.mark.parametrize("name", CASES)
def test_bf16_forward(name):
    g = load_golden(f"model_{name}.npz")
    m, cfg, sd, x = build(g, None)
    ref = O.mask_trans_unet_forward(x, sd, cfg)                         # fp32 CPU oracle, full tensors
    m.eval()
    m.record = {}
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        onehot = m(x.cuda())
    rec, m.record = m.record, None
    logits = from_cl(rec["logits"]).cpu()
    boxes_equal = all(torch.equal(rec[f"box{i}"].cpu(), ref["boxes"][i]) for i in (1, 2, 3))
    err = rel_err(logits, ref["logits"])
    d2s = O.depth_to_space(ref["logits"])
    top2 = d2s.topk(2, dim=1).values
    margin = (top2[:, 0] - top2[:, 1]) > 0.25
    flips = (onehot.cpu().argmax(1) != ref["onehot"].argmax(1))
    flip_all = float(flips.float().mean())
    flip_margin = float(flips[margin].float().mean()) if margin.any() else 0.0
    print(f"\n[bf16 {name}] logits rel err {err:.3e}; argmax flips {flip_all:.3e} (all) {flip_margin:.3e} "
          f"(margin>0.25, {float(margin.float().mean()):.2f} of voxels); boxes equal: {boxes_equal}")
    assert err < 8e-2
    assert flip_all < 3e-2
    assert flip_margin < 1e-2


def test_bf16_with_teacher_forced_boxes_and_tensor_cores_toggle():
    """Teacher-forced ROI boxes remove the data-dependent discontinuity; the tcgen05
    and the CUDA-core convolution paths must agree to bf16 rounding."""
    g = load_golden("model_c2_384x384x16_wellformed.npz")
    m, cfg, sd, x = build(g, "bf16")
    m.eval()
    m.forced_boxes = {i: torch.from_numpy(g[f"box{i}"]) for i in (1, 2, 3)}
    outs = {}
    for tc in (True, False):
        m.use_tensor_cores = tc
        outs[tc] = from_cl(m.forward_logits(x.cuda())).cpu()
    print(f"\n[bf16 tc vs cuda-core] rel diff {rel_err(outs[True], outs[False]):.3e}; "
          f"vs golden {rel_err(sub(outs[True]), g['logits']):.3e}")
    # the tcgen05 path also rounds the conv WEIGHTS to bf16 (the CUDA-core path keeps them fp32), so the
    # two bf16 pipelines differ by about as much as either differs from fp32
    assert rel_err(outs[True], outs[False]) < 8e-2
    assert rel_err(sub(outs[True]), g["logits"]) < 8e-2


def test_module_api_contract():
    """state_dict layout, eval/train return types, determinism, predict_labels, error paths."""
    from lintransunet_b200 import MaskTransUnet, get_model_dict
    cfg = O.UnetConfig(dim_output=3)
    cls = get_model_dict("MaskTransUnet")
    assert cls is MaskTransUnet
    m = cls(num_layers=[16, 32, 64, 128, 256], roi_size_list=[100, 65, 40, 25, 10],
            is_roi_list=[False, True, True, True, True], dim_input=1, dim_output=3)
    spec = {k: s for k, s, _ in O.state_dict_spec(cfg)}
    sd = m.state_dict()
    assert set(sd) == set(spec) and all(tuple(sd[k].shape) == spec[k] for k in sd)
    m.cuda().eval()
    x = O.make_input((2, 1, 32, 64, 8), seed=3).cuda()
    y1, y2 = m(x), m(x)
    assert y1.shape == (2, 3, 32, 64, 8) and y1.dtype == torch.float32 and torch.equal(y1, y2)
    labels = m.predict_labels(x)
    assert labels.dtype == torch.uint8 and torch.equal(labels.long(), y1.argmax(1))
    # weights changed in place -> derived caches are rebuilt
    with torch.no_grad():
        m.decode.final_block.bias[0:4] += 100.0       # class 0 wins everywhere (head channel = c*4+kh*2+kw)
    assert float(m(x)[:, 0].mean()) == 1.0
    m.train()
    with pytest.raises(NotImplementedError):
        m(x)                                   # training (dropout, native backward) exists under autocast only
    with pytest.raises(RuntimeError):
        m.eval()(x.cpu())                      # no CPU fallback
    with pytest.raises(ValueError):
        m(torch.zeros(1, 1, 48, 64, 8, device="cuda"))


def test_cuda_graph_replay_is_bit_identical_to_eager():
    """The captured forward (no host syncs: boxes stay on the device) must equal the eager launch sequence,
    for new inputs of the same shape, for both precisions."""
    g = load_golden("model_c3_64x96x32_b2.npz")
    m, cfg, sd, x = build(g, "bf16")
    m.eval()
    x1 = x.cuda()
    x2 = O.make_input(tuple(x.shape), seed=77, blob=True).cuda()
    for prec in ("bf16", "fp32"):
        m.precision = prec
        m.use_cuda_graphs = False
        e1, e2, l2 = m(x1), m(x2), m.predict_labels(x2).clone()
        m.use_cuda_graphs = True
        g1, g2, g1b = m(x1), m(x2), m(x1)            # capture on x1, replay on x2, replay on x1 again
        assert torch.equal(e1, g1) and torch.equal(e2, g2) and torch.equal(g1, g1b), prec
        assert torch.equal(m.predict_labels(x2), l2)
        assert not torch.equal(e1, e2)


def test_dead_mask_head_does_not_change_the_result():
    """Inference does not launch the mask head of a level without a ROI bridge (nobody reads it: the eval result is
    the argmax of the final block, model/trans_3DUnet.py:199-201): logits, one-hot and labels are bit-identical to
    the forward that computes it, in both precisions; training mode always computes it (deep supervision)."""
    g = load_golden("model_c3_64x96x32_b2.npz")
    m, cfg, sd, x = build(g, "bf16")
    m.eval()
    xc = x.cuda()
    for prec in ("bf16", "fp32"):
        m.precision = prec
        m.skip_dead_mask_head = False
        full = (m.forward_logits(xc), m(xc), m.predict_labels(xc).clone())
        m.skip_dead_mask_head = True
        lean = (m.forward_logits(xc), m(xc), m.predict_labels(xc).clone())
        for a, b in zip(full, lean):
            assert torch.equal(a, b), prec
    m.train()
    m.precision = "bf16"
    with torch.no_grad():
        probs, mask_list = m(xc)
    assert len(mask_list) == 4 and mask_list[-1].shape[2:] == (32, 48, 32)


def test_result_does_not_depend_on_batch_composition():
    """Every reduction is split by per-sample rules only (never by the batch size), so a patch gives the
    same bits alone, in a batch of 3 or in a batch of 5 -- the property that makes the N-GPU sliding window
    bit-identical to the 1-GPU one (tools/mgpu_check.py checks it across ranks)."""
    g = load_golden("model_c3_64x96x32_b2.npz")
    m, cfg, sd, x = build(g, "bf16")
    m.eval()
    xs = torch.cat([O.make_input((1, 1, 64, 96, 32), seed=90 + i, blob=bool(i % 2)) for i in range(5)]).cuda()
    for prec in ("bf16", "fp32"):
        m.precision = prec
        full = m.forward_logits(xs)
        for lo, hi in ((0, 1), (1, 4), (4, 5), (2, 3)):
            part = m.forward_logits(xs[lo:hi].contiguous())
            assert torch.equal(part, full[lo:hi]), (prec, lo, hi)
