"""GPU: the post-processing kernels behind the inference driver (ltu_vote_decide, ltu_keep_largest_component,
ltu_overlap_counts; exact integer / byte work) against oracle/postproc.py and the reference-generated golden
metric values, and the driver end to end on synthetic cases."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import postproc as PP
from tests.helpers import load_golden

pytestmark = pytest.mark.gpu


def _votes(shape, C, seed, max_n=8):
    """Random vote volumes with coverage counts n in 1..max_n (including n = 3, 6, ... of overlap 0.6)."""
    g = np.random.default_rng(seed)
    n = g.integers(1, max_n + 1, size=shape)
    v = np.zeros((C,) + tuple(shape), dtype=np.uint8)
    left = n.copy()
    for c in range(C - 1):
        k = g.integers(0, left + 1)
        v[c] = k
        left = left - k
    v[C - 1] = left
    return v


@pytest.mark.parametrize("shape,C", [((16, 12, 8), 3), ((7, 5, 3), 2), ((33, 20, 12), 4)])
def test_vote_decide_and_fractions_are_bit_exact(shape, C):
    from lintransunet_b200 import ops
    v = _votes(shape, C, seed=sum(shape) + C)
    frac = PP.vote_fractions(v)
    vd = torch.from_numpy(v).cuda()
    assert np.array_equal(ops.vote_fractions(vd).cpu().numpy(), frac)                   # IEEE division, every n
    for thr in (0.5, 0.3, 0.75):
        got = ops.vote_decide(vd, ops.DECIDE_THRESHOLD, thr).cpu().numpy()
        assert np.array_equal(got, PP.decide_threshold(frac, thr)), thr
    got = ops.vote_decide(vd, ops.DECIDE_ROUND).cpu().numpy()
    assert np.array_equal(got, PP.decide_round(frac))
    assert ((frac == 0.5) & (got == 0)).any() or C > 3                                  # half-to-even cases are present


def _random_onehot(shape, C, density, seed, blobs=0):
    g = np.random.default_rng(seed)
    lab = np.zeros(shape, dtype=np.uint8)
    noise = g.random(shape) < density
    lab[noise] = g.integers(1, C, size=int(noise.sum()))
    for _ in range(blobs):
        h0, w0, d0 = (int(g.integers(0, s - 2)) for s in shape)
        dh, dw, dd = (int(g.integers(2, max(3, s // 3))) for s in shape)
        lab[h0:h0 + dh, w0:w0 + dw, d0:d0 + dd] = int(g.integers(1, C))
    return np.stack([(lab == c) for c in range(C)]).astype(np.uint8)


@pytest.mark.parametrize("connectivity", [1, 2, 3])
@pytest.mark.parametrize("shape,density,blobs", [((12, 10, 8), 0.15, 0), ((24, 20, 16), 0.05, 4), ((9, 7, 5), 0.5, 0),
                                                 ((40, 33, 21), 0.25, 6), ((64, 64, 32), 0.02, 12)])
def test_keep_largest_component_matches_restated_monai(shape, density, blobs, connectivity):
    from lintransunet_b200 import ops
    oh = _random_onehot(shape, 3, density, seed=shape[0] * 7 + connectivity, blobs=blobs)
    for independent in (False, True):
        want = PP.keep_largest_connected_component(oh, [1, 2], independent=independent, connectivity=connectivity)
        got = ops.keep_largest_component_(torch.from_numpy(oh.copy()).cuda(), [1, 2], connectivity=connectivity,
                                          independent=independent).cpu().numpy()
        assert np.array_equal(got, want), (independent, int((got != want).sum()))
    # a single applied label (the commented-out binary variant, inference_embed_attn.py:109)
    want = PP.keep_largest_connected_component(oh, [1], connectivity=connectivity)
    got = ops.keep_largest_component_(torch.from_numpy(oh.copy()).cuda(), [1], connectivity=connectivity).cpu().numpy()
    assert np.array_equal(got, want)


def test_keep_largest_component_ties_empty_and_full():
    from lintransunet_b200 import ops
    oh = np.zeros((3, 6, 6, 4), dtype=np.uint8)
    oh[1, 0:2, 0:2, 0] = 1
    oh[2, 2, 2, 1] = 1
    oh[2, 4:6, 4:6, 3] = 1
    oh[0] = 1 - oh[1] - oh[2]
    for conn in (1, 3):                                          # conn 1: 4-4 tie, first block in raster order wins
        want = PP.keep_largest_connected_component(oh, [1, 2], connectivity=conn)
        got = ops.keep_largest_component_(torch.from_numpy(oh.copy()).cuda(), [1, 2], connectivity=conn).cpu().numpy()
        assert np.array_equal(got, want)
    z = torch.zeros(3, 8, 8, 4, dtype=torch.uint8, device="cuda")
    assert int(ops.keep_largest_component_(z, [1, 2]).sum()) == 0
    full = torch.zeros(2, 16, 16, 8, dtype=torch.uint8, device="cuda")
    full[1] = 1                                                  # one component of every voxel (long union chains)
    assert int(ops.keep_largest_component_(full, [1]).sum()) == 16 * 16 * 8
    # run to run identical (the union-find result does not depend on thread timing)
    oh = _random_onehot((48, 40, 24), 3, 0.3, seed=11, blobs=5)
    a = ops.keep_largest_component_(torch.from_numpy(oh.copy()).cuda(), [1, 2]).cpu()
    b = ops.keep_largest_component_(torch.from_numpy(oh.copy()).cuda(), [1, 2]).cpu()
    assert torch.equal(a, b)


@pytest.mark.parametrize("name", ["b0", "b1", "b_empty_pred", "b_empty_target", "m0", "m1", "m_empty_pred", "m_empty_target"])
def test_overlap_counts_and_metrics_match_reference_values(name):
    from lintransunet_b200 import ops
    from lintransunet_b200.inference import metrics_from_counts
    g = load_golden("postproc.npz")
    multi = name.startswith("m")
    C = 3 if multi else 2
    pred = np.stack([(g[f"{name}_pred"] == c) for c in range(C)]).astype(np.uint8)
    counts = ops.overlap_counts(torch.from_numpy(pred).cuda(), torch.from_numpy(g[f"{name}_target"]).cuda())
    assert np.array_equal(counts.cpu().numpy(), PP.overlap_counts(pred, g[f"{name}_target"]))
    np.testing.assert_allclose(list(metrics_from_counts(counts, multi).values()), g[f"{name}_values"], rtol=0, atol=2e-6)


def _model(dim_output, seed=0):
    from lintransunet_b200 import MaskTransUnet
    torch.manual_seed(seed)
    return MaskTransUnet([16, 32, 64, 128, 256], [100, 65, 40, 25, 10], [False, True, True, True, True], 1,
                         dim_output).cuda().eval()


@pytest.mark.parametrize("dim_output", [2, 3])
def test_segment_volume_equals_the_reference_recipe_on_the_same_votes(dim_output):
    """segment_volume = sliding window -> decision -> (largest component -> background) on the GPU; the same recipe
    restated on the CPU from the same stitched fractions gives the same bytes, and the saved array has the
    reference's layout and dtype."""
    from lintransunet_b200.inference import prediction_array, segment_volume
    from lintransunet_b200.sliding_window import sliding_window_inference
    m = _model(dim_output)
    vol = torch.randn(1, 1, 96, 64, 40, generator=torch.Generator().manual_seed(3)).pin_memory()
    roi, ov = (64, 64, 16), 0.6                                  # interval int(16*0.4) = 6 along D: coverage counts up to 3
    multi = dim_output > 2
    onehot = segment_volume(m, vol, roi, sw_batch_size=4, overlap=ov)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        frac = sliding_window_inference(vol.cuda(), roi, 4, m, overlap=ov, sigma_scale=0)[0].cpu().numpy()
    if multi:
        want = PP.decide_round(frac)
        want = PP.keep_largest_connected_component(want, [1, 2], independent=False, connectivity=3)
        want = PP.background_from_rest(want).astype(np.uint8)
        saved = np.argmax(want, 0).transpose(2, 0, 1).astype(np.int64)
    else:
        want = PP.decide_threshold(frac, 0.5)
        saved = want[1].astype(np.float32).transpose(2, 0, 1)
    assert onehot.dtype == torch.uint8 and np.array_equal(onehot.cpu().numpy(), want)
    arr = prediction_array(onehot, multi)
    assert arr.dtype == saved.dtype and arr.shape == (40, 96, 64) and np.array_equal(arr, saved)


def test_driver_main_writes_predictions_and_summary(tmp_path):
    from lintransunet_b200 import inference
    g = np.random.default_rng(5)
    data = tmp_path / "data"
    (data / "image").mkdir(parents=True)
    (data / "label").mkdir()
    for i in range(2):
        np.save(data / "image" / f"case{i}.npy", g.integers(-200, 400, size=(24, 64, 64)).astype(np.int16))   # [D,H,W]
        np.save(data / "label" / f"case{i}.npy", g.integers(0, 3, size=(24, 64, 64)).astype(np.uint8))
    out = tmp_path / "pred"
    summary = inference.main(["--dir_data", str(data), "--dim_output", "3", "--roi_size", "64", "--depth_size", "16",
                              "--is_save", "--saved_folder", str(out), "--summary_json", str(tmp_path / "s.json")])
    files = sorted(os.listdir(out))
    assert files == ["case0.npy_multi.npy", "case1.npy_multi.npy"]         # '{:0>4}'.format(name) + '_multi' (:162)
    arr = np.load(out / files[0])
    assert arr.shape == (24, 64, 64) and arr.dtype == np.int64 and arr.min() >= 0 and arr.max() <= 2
    keys = ["DiceClassLoss0", "DiceClassLoss", "DiceClassLoss2", "Recall", "Precision", "Recall2", "Precision2",
            "LocalizationLoss"]
    assert summary["criterions"] == keys and len(summary["patient_0"]) == 2
    assert json.load(open(tmp_path / "s.json"))["summary_0"] == summary["summary_0"]
    assert all(np.isfinite(v) for v in summary["summary_0"])
