"""CPU: pin oracle/ltu_oracle.py against vectors produced by the unmodified reference
(tools/make_golden.py).  fp32 tolerance 5e-5 (max|a-b|/max|ref|): both sides are fp32
CPU programs that differ only in summation order."""
import numpy as np
import pytest
import torch

from oracle import ltu_oracle as O
from tests.helpers import load_golden, rel_err, sub

TOL = 5e-5


def test_state_dict_layout_is_the_reference_one():
    cfg = O.UnetConfig(dim_output=2)
    spec = O.state_dict_spec(cfg)
    assert len(spec) == 614                               # 614 keys
    assert sum(int(np.prod(s)) for _, s, _ in spec) == 20872836
    keys = {k for k, _, _ in spec}
    for k in ("encode.input_block.weight",
              "decode.bridge_list.1.transformer.down_embed.module_list.0.0.weight",
              "decode.bridge_list.3.transformer.up_embed.module_list.0.1.bias",
              "decode.bridge_list.4.transformer.pos_encoders.7.proj.weight",
              "decode.bridge_list.2.transformer.layers.7.self_attn.linears.3.bias",
              "decode.att_conv_list.0.psi.0.weight", "decode.final_block.bias"):
        assert k in keys


def test_attention_core():
    g = load_golden("ops.npz")
    for tag in ("a", "b"):
        q, k, v = (torch.from_numpy(g[f"attn_{tag}_{n}"]) for n in "qkv")
        out = O.efficient_attention(q, k, v)
        assert rel_err(out, g[f"attn_{tag}_out"]) < TOL


def test_encoder_layer_and_posconv():
    g = load_golden("ops.npz")
    sd = O.make_state_dict(O.UnetConfig(), seed=3)
    y = O.encoder_layer(torch.from_numpy(g["layer_x"]), sd,
                        "decode.bridge_list.1.transformer.layers.2", nhead=4)
    assert rel_err(y, g["layer_out"]) < TOL
    p = "decode.bridge_list.1.transformer.pos_encoder.proj"
    z = O.pos_embedding(torch.from_numpy(g["pos_x"]), sd[p + ".weight"], sd[p + ".bias"])
    assert rel_err(z, g["pos_out"]) < TOL


def test_fisheye_maps_bit_exact():
    g = load_golden("ops.npz")
    for i, (x0, x1, h, roi, eroi) in enumerate(g["fish_cases"]):
        a = torch.tensor([[x0]], dtype=torch.float32)
        b = torch.tensor([[x1]], dtype=torch.float32)
        f = O.fisheye_forward_coords(a, b, int(h), int(roi), int(eroi))[0].numpy()
        r = O.fisheye_back_coords(a, b, int(h), int(roi), int(eroi))[0].numpy()
        assert np.array_equal(f, g[f"fish_fwd{i}"], equal_nan=True), i
        assert np.array_equal(r, g[f"fish_back{i}"], equal_nan=True), i


def _golden_masks(g):
    shape = tuple(int(s) for s in g["box_masks_shape"])
    bits = np.unpackbits(g["box_masks_packed"])[: int(np.prod(shape))]
    return torch.from_numpy(bits.reshape(shape).astype(np.float32))


def test_roi_boxes_exact_and_resample():
    g = load_golden("ops.npz")
    masks = _golden_masks(g)
    rc = O.UnetConfig().roi_consts(1)
    box = O.roi_boxes(masks, rc["min_h"], rc["min_w"])
    assert np.array_equal(box.numpy(), g["box_out"])
    feat = torch.randn(5, 4, 96, 96, 8, generator=torch.Generator().manual_seed(int(g["resample_feat_seed"])))
    x0, y0, x1, y1 = box[:, 0:1], box[:, 1:2], box[:, 3:4], box[:, 4:5]
    roi = O.separable_resample(feat, O.fisheye_forward_coords(x0, x1, 95, rc["h_roi"], rc["eval_h"]),
                               O.fisheye_forward_coords(y0, y1, 95, rc["w_roi"], rc["eval_w"]))
    assert rel_err(sub(roi, 65536), g["resample_roi"]) < TOL
    back = O.separable_resample(roi, O.fisheye_back_coords(x0, x1, 95, rc["h_roi"], rc["eval_h"]),
                                O.fisheye_back_coords(y0, y1, 95, rc["w_roi"], rc["eval_w"]))
    assert rel_err(sub(back, 65536), g["resample_back"]) < TOL


@pytest.mark.parametrize("name", ["c2_64x64x16", "c3_64x96x32_b2", "c2_384x384x16_wellformed"])
def test_whole_model(name):
    g = load_golden(f"model_{name}.npz")
    cfg = O.UnetConfig(dim_output=int(g["dim_output"]))
    sd = O.make_state_dict(cfg, seed=int(g["seed_w"]))
    x = O.make_input(tuple(int(s) for s in g["shape"]), seed=int(g["seed_x"]), blob=bool(g["blob"]))
    o = O.mask_trans_unet_forward(x, sd, cfg, want_taps=True)
    for i in (1, 2, 3):
        assert np.array_equal(o["boxes"][i].numpy(), g[f"box{i}"]), f"box {i}"
    t = o["taps"]
    assert rel_err(sub(t["encode.bottle"]), g["bottle"]) < TOL
    for i in range(4):
        assert rel_err(sub(t[f"encode.skip{i}"]), g[f"skip{i}"]) < TOL
        assert rel_err(sub(t[f"decode.up{i}"]), g[f"up{i}"]) < TOL
        assert rel_err(sub(o["mask_list"][i]), g[f"mask{i}"]) < TOL
    assert rel_err(sub(t["bridge_bottle"]), g["bridge4"]) < TOL
    for i in (1, 2, 3):
        assert rel_err(sub(t[f"bridged_skip{i}"]), g[f"bridge{i}"]) < 2e-4
    assert rel_err(sub(o["logits"]), g["logits"]) < TOL
    assert rel_err(sub(o["probs"]), g["probs"]) < TOL
    assert float(np.mean(sub(o["onehot"]) != g["onehot"])) <= 1e-4
