"""CPU: the gradient oracle for the backward path is pinned to the reference.
oracle/train_step.py (loss recipe) + autograd through oracle/ltu_oracle.py must reproduce the loss terms and
the parameter gradients the unmodified reference produced (tests/golden/train_c2_64x64x16.npz,
tools/make_golden_train.py)."""
import numpy as np
import torch

from oracle import ltu_oracle as O
from oracle import train_step as T
from tests.helpers import load_golden


def _gsub(t, n=64):
    f = t.detach().reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step][:n].to(torch.float32).numpy().copy()


def test_oracle_train_step_matches_reference_loss_and_gradients():
    g = load_golden("train_c2_64x64x16.npz")
    cfg = O.UnetConfig(dim_output=2)
    sd = {k: v.clone().requires_grad_(True) for k, v in O.make_state_dict(cfg, seed=0).items()}
    x = O.make_input((1, 1, 64, 64, 16), seed=1, blob=True)
    masks = torch.from_numpy(g["masks"]).long()
    out = O.mask_trans_unet_forward(x, sd, cfg)
    total, terms = T.train_loss(out["probs"], out["mask_list"], masks)
    got_terms = np.asarray([[float(v.detach()) for v in row] for row in terms])
    np.testing.assert_allclose(got_terms, g["terms"], rtol=2e-5, atol=2e-6)
    assert abs(float(total) - float(g["total"])) <= 2e-5 * abs(float(g["total"]))
    total.backward()
    names = [str(n) for n in g["grad_names"]]
    assert sorted(n for n, p in sd.items() if p.grad is not None) == sorted(names)            # same live parameters
    assert sorted(n for n, p in sd.items() if p.grad is None) == sorted(str(n) for n in g["dead_names"])
    # Gradients below 1e-5 of the largest one are cancellation noise in the reference itself (conv biases in front of
    # an InstanceNorm, the K-projection bias under the softmax over tokens, layer 0 of the ROI bridges whose input
    # is a near-constant channel): there both sides only have to be negligible.  Everything else is
    # compared tightly.
    ref_norm = dict(zip(names, (float(v) for v in g["grad_norms"])))
    floor = 1e-5 * max(ref_norm.values())
    worst_norm, worst_elem, negligible = 0.0, 0.0, 0
    for name in names:
        grad = sd[name].grad
        norm = float(grad.double().norm())
        if ref_norm[name] < floor:
            negligible += 1
            assert norm < 2 * floor, name
            continue
        worst_norm = max(worst_norm, abs(norm - ref_norm[name]) / ref_norm[name])
        ref = g["g:" + name]
        worst_elem = max(worst_elem, float(np.abs(_gsub(grad) - ref).max()) / float(np.abs(ref).max()))
    print(f"\n[train oracle] {len(names)} gradients, {len(names) - negligible} above the noise floor: worst relative norm "
          f"error {worst_norm:.2e}, worst element error / max|g| {worst_elem:.2e}")
    assert worst_norm <= 1e-3 and worst_elem <= 2e-3 and negligible < len(names) // 2
