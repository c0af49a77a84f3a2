#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference, build container only) on seeded synthetic weights and inputs.

The reference cannot travel to the GPU box, the vectors can: they pin
oracle/ltu_oracle.py (tests/test_oracle_golden.py) and are a second, independent
check for the CUDA path (tests/test_model_gpu.py).  Weights come from
oracle.ltu_oracle.make_state_dict (names/shapes are the interface; the reference
consumes them through load_state_dict(strict=True)), so no RNG init stream has to
match.  Large tensors are stored as a deterministic strided subsample (`sub`).

Usage:  python tools/make_golden.py [--ref /root/reference]
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ltu_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def sub(t: torch.Tensor, n: int = 16384) -> np.ndarray:
    """Deterministic strided subsample used on both sides of every comparison."""
    f = t.detach().reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step].to(torch.float32).cpu().numpy().copy()


def model_case(ref_mod, name, shape, dim_output, seed_w, seed_x, blob):
    cfg = O.UnetConfig(dim_output=dim_output)
    sd = O.make_state_dict(cfg, seed=seed_w)
    m = ref_mod.get_model_dict("MaskTransUnet")(
        num_layers=list(cfg.num_layers), roi_size_list=list(cfg.roi_size_list),
        is_roi_list=list(cfg.is_roi_list), dim_input=1, dim_output=dim_output, dropout=0.0)
    m.load_state_dict(sd, strict=True)
    x = O.make_input(shape, seed=seed_x, blob=blob)
    taps = {}
    hooks = []

    def grab(key):
        def fn(mod, inp, out):
            taps[key] = out[0] if isinstance(out, tuple) and key != "encode" else out
        return fn

    hooks.append(m.decode.final_block.register_forward_hook(grab("logits")))
    hooks.append(m.encode.register_forward_hook(grab("encode")))
    for i in range(1, 5):
        hooks.append(m.decode.bridge_list[i].register_forward_hook(grab(f"bridge{i}")))
    for i in range(4):
        hooks.append(m.decode.block_list[i].register_forward_hook(grab(f"up{i}")))
    boxes = {}
    for i in (1, 2, 3):
        br = m.decode.bridge_list[i]
        orig = br.get_mask_boundary2

        def wrapped(mask, _o=orig, _i=i):
            b = _o(mask)
            boxes[_i] = b.clone()
            return b
        br.get_mask_boundary2 = wrapped
    with torch.no_grad():
        m.train()                       # dropout=0.0 => deterministic; returns (probs, mask_list)
        probs, mask_list = m(x)
        m.eval()
        onehot = m(x)
    for h in hooks:
        h.remove()
    bottle, skips = taps["encode"]
    out = dict(
        shape=np.array(shape), dim_output=np.array(dim_output), seed_w=np.array(seed_w),
        seed_x=np.array(seed_x), blob=np.array(int(blob)),
        logits=sub(taps["logits"]), probs=sub(probs), onehot=sub(onehot),
        logits_absmax=np.array(float(taps["logits"].abs().max())),
        argmax_hist=np.bincount(probs.argmax(1).reshape(-1).numpy(), minlength=dim_output),
        bottle=sub(bottle),
    )
    for i, s in enumerate(skips):
        out[f"skip{i}"] = sub(s)
    for i in range(1, 5):
        out[f"bridge{i}"] = sub(taps[f"bridge{i}"])
    for i in range(4):
        out[f"up{i}"] = sub(taps[f"up{i}"])
        out[f"mask{i}"] = sub(mask_list[i])
    for i, b in boxes.items():
        out[f"box{i}"] = b.numpy()
    path = os.path.join(GOLD, f"model_{name}.npz")
    np.savez_compressed(path, **out)
    # cross-check the oracle right away (the pytest does the same from the file)
    o = O.mask_trans_unet_forward(x, sd, cfg)
    err = float((o["logits"] - taps["logits"]).abs().max() / taps["logits"].abs().max())
    mism = float((o["onehot"] != onehot).float().mean())
    print(f"[{name}] boxes={ {k: v.tolist() for k, v in boxes.items()} }")
    print(f"[{name}] oracle vs reference: logits rel err {err:.3e}, onehot mismatch {mism:.3e}, "
          f"file {os.path.getsize(path)/1e6:.2f} MB")


def op_cases(ref_mod):
    import model.trans_block as TB
    import model.Unet_3Dblock as UB
    g = torch.Generator().manual_seed(7)
    out = {}
    # --- a1: linear_attention core, ragged N, h in {4, 8}
    for tag, (B, h, N) in {"a": (2, 4, 100), "b": (1, 8, 333)}.items():
        q, k, v = (torch.randn(B, h, N, 32, generator=g) * 2 for _ in range(3))
        o, _ = TB.linear_attention(q, k, v, dropout=torch.nn.Dropout(0.0))
        out.update({f"attn_{tag}_q": q.numpy(), f"attn_{tag}_k": k.numpy(),
                    f"attn_{tag}_v": v.numpy(), f"attn_{tag}_out": o.numpy()})
    # --- a3: one encoder layer, d_model 128 / 4 heads
    cfg = O.UnetConfig()
    sd = O.make_state_dict(cfg, seed=3)
    pre = "decode.bridge_list.1.transformer.layers.2"
    layer = TB.SelfAttentionLayer(d_model=128, nhead=4, dim_feedforward=256, dropout=0.0)
    layer.load_state_dict({k[len(pre) + 1:]: v for k, v in sd.items() if k.startswith(pre + ".")})
    layer.eval()
    x = torch.randn(2, 77, 128, generator=g)
    with torch.no_grad():
        out["layer_x"] = x.numpy()
        out["layer_out"] = layer(x).numpy()
    # --- a4: positional conv on the permuted view, exactly like EmbedAttention3DBlock does
    pe = TB.Conv3dPosEmbedding(dim=128, dropout=0.0)
    pw = sd["decode.bridge_list.1.transformer.pos_encoder.proj.weight"]
    pb = sd["decode.bridge_list.1.transformer.pos_encoder.proj.bias"]
    pe.load_state_dict({"proj.weight": pw, "proj.bias": pb})
    pe.eval()
    vol = torch.randn(1, 128, 5, 4, 6, generator=g)            # [B,C,H,W,D]
    with torch.no_grad():
        y = pe(vol.permute(0, 1, 4, 2, 3)).permute(0, 1, 3, 4, 2)
    out["pos_x"] = vol.numpy()
    out["pos_out"] = y.numpy()
    # --- a9: fisheye index maps (well-formed and degenerate boxes)
    cases = [(7.5, 16.5, 23, 25, 30), (11.0, 35.0, 47, 40, 48), (17.5, 74.5, 95, 65, 78),
             (6.5, -4.5, 3, 25, 30), (19.5, 12.5, 31, 65, 78), (0.0, 31.0, 31, 39, 46)]
    fw, bw = [], []
    for (x0, x1, h, roi, eroi) in cases:
        a = torch.tensor([[x0]], dtype=torch.float32)
        b = torch.tensor([[x1]], dtype=torch.float32)
        fw.append(UB.get_transfer_index(a, b, h, roi, eroi, device="cpu")[0].numpy())
        bw.append(UB.get_transfer_back_index(a, b, h, roi, eroi, device="cpu")[0].numpy())
    out["fish_cases"] = np.array(cases, dtype=np.float64)
    for i, (a, b) in enumerate(zip(fw, bw)):
        out[f"fish_fwd{i}"] = a
        out[f"fish_back{i}"] = b
    # --- a8: ROI boxes from masks (blob, empty, full, tiny) at a well-formed plane size
    br = UB.ROIBridge(in_dim=32, d_model=128, nhead=4, dropout=0.0, N=1, roi_size=65)
    masks = torch.zeros(5, 1, 96, 96, 8)
    hh = torch.arange(96).view(96, 1, 1)
    ww = torch.arange(96).view(1, 96, 1)
    masks[0, 0] = (((hh - 40) / 22.0) ** 2 + ((ww - 55) / 15.0) ** 2 < 1).float().expand(96, 96, 8) * 0.9
    masks[2, 0] = 1.0
    masks[3, 0, 50:53, 20:22, 3] = 0.7
    masks[4, 0] = torch.rand(96, 96, 8, generator=g)
    with torch.no_grad():
        out["box_masks_packed"] = np.packbits((masks >= 0.5).numpy())
        out["box_masks_shape"] = np.array(masks.shape)
        out["box_out"] = br.get_mask_boundary2(masks >= 0.5).numpy()
        # --- a7/a9: resample there and back with those boxes
        feat = torch.randn(5, 4, 96, 96, 8, generator=torch.Generator().manual_seed(11))
        roi = br.roi_alignment2(feat, torch.from_numpy(out["box_out"]))
        back = br.post_processing2(feat, roi, torch.from_numpy(out["box_out"]))
    out["resample_feat_seed"] = np.array(11)          # feat is regenerated from this seed by the tests
    out["resample_roi"] = sub(roi, 65536)
    out["resample_back"] = sub(back, 65536)
    path = os.path.join(GOLD, "ops.npz")
    np.savez_compressed(path, **out)
    print(f"[ops] wrote {path} ({os.path.getsize(path)/1e6:.2f} MB)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    sys.path.insert(0, args.ref)
    import model.trans_3DUnet as ref_mod            # the unmodified reference
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    op_cases(ref_mod)
    model_case(ref_mod, "c2_64x64x16", (1, 1, 64, 64, 16), 2, 0, 1, False)
    model_case(ref_mod, "c3_64x96x32_b2", (2, 1, 64, 96, 32), 3, 5, 6, True)
    model_case(ref_mod, "c2_384x384x16_wellformed", (1, 1, 384, 384, 16), 2, 0, 1, True)


if __name__ == "__main__":
    main()
