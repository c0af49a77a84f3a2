"""CPU tests of the post-processing oracle (oracle/postproc.py) and of the host logic of the inference driver
(lintransunet_b200/inference.py): pinned to the reference's own metric classes through tests/golden/postproc.npz
(tools/make_golden_postproc.py)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import postproc as PP
from tests.helpers import load_golden

BINARY = ["DiceClassLoss", "Recall", "Precision", "LocalizationLoss"]
MULTI = ["DiceClassLoss0", "DiceClassLoss", "DiceClassLoss2", "Recall", "Precision", "Recall2", "Precision2",
         "LocalizationLoss"]
CASES_B = ["b0", "b1", "b_empty_pred", "b_empty_target"]
CASES_M = ["m0", "m1", "m_empty_pred", "m_empty_target"]
TOL = 2e-6          # the reference evaluates in fp32


def _onehot(lab: np.ndarray, C: int) -> torch.Tensor:
    return F.one_hot(torch.from_numpy(lab).long(), C).permute(3, 0, 1, 2)[None].float()


@pytest.mark.parametrize("name", CASES_B)
def test_oracle_binary_metrics_match_reference(name):
    g = load_golden("postproc.npz")
    predict = _onehot(g[f"{name}_pred"], 2)
    masks = torch.from_numpy(g[f"{name}_target"]).long()[None, None]
    got = [float(PP.dice_class_loss(predict, masks)), float(PP.recall(predict, masks)),
           float(PP.precision(predict, masks)), float(PP.localization_loss(predict, masks))]
    np.testing.assert_allclose(got, g[f"{name}_values"], rtol=0, atol=TOL)


@pytest.mark.parametrize("name", CASES_M)
def test_oracle_multi_metrics_match_reference(name):
    g = load_golden("postproc.npz")
    predict, label = _onehot(g[f"{name}_pred"], 3), _onehot(g[f"{name}_target"], 3)
    got = [float(PP.dice_class_loss_multi(predict, label, 0)), float(PP.dice_class_loss_multi(predict, label, 1)),
           float(PP.dice_class_loss_multi(predict, label, 2)), float(PP.recall_multi(predict, label, 1)),
           float(PP.precision_multi(predict, label, 1)), float(PP.recall_multi(predict, label, 2)),
           float(PP.precision_multi(predict, label, 2)), float(PP.localization_loss_multi(predict, label))]
    np.testing.assert_allclose(got, g[f"{name}_values"], rtol=0, atol=TOL)


@pytest.mark.parametrize("name", CASES_B + CASES_M)
def test_driver_metrics_from_counts_match_reference(name):
    """Host side of the driver: integer counts (here from the numpy oracle, on the GPU from ltu_overlap_counts)
    -> the reference's printed values, under the reference's criterion names and in its order."""
    from lintransunet_b200.inference import metrics_from_counts
    g = load_golden("postproc.npz")
    multi = name.startswith("m")
    C = 3 if multi else 2
    pred = np.stack([(g[f"{name}_pred"] == c) for c in range(C)]).astype(np.uint8)
    counts = torch.from_numpy(PP.overlap_counts(pred, g[f"{name}_target"]))
    m = metrics_from_counts(counts, multi)
    assert list(m.keys()) == (MULTI if multi else BINARY)
    np.testing.assert_allclose(list(m.values()), g[f"{name}_values"], rtol=0, atol=TOL)


def test_decisions_on_vote_fractions():
    votes = np.array([[1, 2, 3, 0, 4, 1], [1, 1, 3, 8, 4, 2], [0, 1, 0, 0, 0, 3]], dtype=np.uint8)   # n = 2,4,6,8,8,6
    frac = PP.vote_fractions(votes)
    assert frac.dtype == np.float32 and np.allclose(frac.sum(0), 1)
    thr = PP.decide_threshold(frac)
    rnd = PP.decide_round(frac)
    # column 0: 1/2 and 1/2 -> both pass `>= 0.5`, neither survives round-half-to-even
    assert thr[:, 0].tolist() == [1, 1, 0] and rnd[:, 0].tolist() == [0, 0, 0]
    # column 2: 3/6 = 0.5 exactly in fp32
    assert frac[0, 2] == np.float32(0.5) and rnd[:, 2].tolist() == [0, 0, 0]
    assert thr[:, 3].tolist() == [0, 1, 0] and rnd[:, 3].tolist() == [0, 1, 0]
    assert rnd[:, 5].tolist() == [0, 0, 0] and thr[:, 5].tolist() == [0, 0, 1]
    assert rnd.sum(0).max() <= 1                       # at most one class can round to 1


def test_keep_largest_component_semantics():
    oh = np.zeros((3, 6, 6, 4), dtype=np.uint8)
    oh[1, 0:2, 0:2, 0] = 1                              # 4 voxels of class 1 ...
    oh[2, 2, 2, 1] = 1                                  # ... touching one voxel of class 2 only through a corner
    oh[2, 4:6, 4:6, 3] = 1                              # 4 voxels, far away
    oh[0] = 1 - oh[1] - oh[2]
    # connectivity 3: the corner contact merges {class-1 block, class-2 voxel} into a component of 5 > 4
    out = PP.keep_largest_connected_component(oh, [1, 2], independent=False, connectivity=3)
    assert out[1].sum() == 4 and out[2].sum() == 1 and out[2, 2, 2, 1] == 1
    assert (out[0] == oh[0]).all()                      # channel 0 is not an applied label
    # connectivity 1: three components of 4, 1, 4 voxels -> tie between the two blocks, the first in raster order wins
    out = PP.keep_largest_connected_component(oh, [1, 2], independent=False, connectivity=1)
    assert out[1].sum() == 4 and out[2].sum() == 0
    # independent: every label keeps its own largest component
    out = PP.keep_largest_connected_component(oh, [1, 2], independent=True, connectivity=1)
    assert out[1].sum() == 4 and out[2].sum() == 4 and out[2, 2, 2, 1] == 0
    # no foreground at all: unchanged
    z = np.zeros_like(oh)
    assert (PP.keep_largest_connected_component(z, [1, 2]) == 0).all()
    fixed = PP.background_from_rest(out)
    assert (fixed.sum(0) == 1).all()


def test_prepare_ct_follows_the_dataset_statements():
    from lintransunet_b200.inference import CT_NORM, prepare_ct
    g = np.random.default_rng(0)
    for dtype in (np.int16, np.float32, np.float64):
        vol = g.integers(-400, 600, size=(5, 8, 6)).astype(dtype)            # stored [D,H,W]
        for multi in (False, True):
            n = CT_NORM[multi]
            img = vol.copy()                                                   # dataset/CT_pancreas_ids.py:219-225
            img[img < n["low_clip"]] = n["low_clip"]
            img[img > n["high_clip"]] = n["high_clip"]
            img = ((img - n["mean"]) / n["std"]).astype(np.float32)
            want = torch.from_numpy(img)[None].permute(0, 2, 3, 1)[None]       # AddChannel, permute(0,2,3,1), batch of 1
            got = prepare_ct(vol, multi)
            assert got.shape == (1, 1, 8, 6, 5) and got.dtype == torch.float32
            assert torch.equal(got, want)
        assert vol.min() < -96                                                  # the input itself is not modified


def test_parser_defaults_are_the_reference_defaults():
    from lintransunet_b200.inference import get_parse
    a = get_parse(["--dir_data", "/tmp/x"])
    assert a.num_layers == [16, 32, 64, 128, 256] and a.roi_size_list == [100, 65, 40, 25, 10]
    assert a.is_roi_list == [False, True, True, True, True] and a.depth_size == 32 and a.dim_output == 2
    assert (a.roi_size, a.sw_batch_size, a.overlap, a.threshold) == (512, 4, 0.6, 0.5)
    b = get_parse(["--dir_data", "/tmp/x", "--num_layers", "8,16,32", "--is_roi_list", "[False, True, True]"])
    assert b.num_layers == [8, 16, 32] and b.is_roi_list == [False, True, True]


def test_integer_form_of_the_half_decisions_is_exact():
    """postproc_kernels.cu decides `fp32(k/n) >= 0.5` as 2k >= n and `rint(fp32(k/n)) == 1` as 2k > n (packed-byte SIMD):
    exhaustive check over every vote count a uint8 volume can hold."""
    n = np.arange(1, 256, dtype=np.int64)[None, :]
    k = np.arange(0, 256, dtype=np.int64)[:, None]
    ok = k <= n
    frac = (k.astype(np.float32) / n.astype(np.float32)).astype(np.float32)
    assert np.array_equal((frac >= np.float32(0.5))[ok], (2 * k >= n)[ok])
    rounded = torch.round(torch.from_numpy(frac)).numpy()
    assert np.array_equal((rounded == 1)[ok], (2 * k > n)[ok])


def _run_based_components(fg: np.ndarray) -> np.ndarray:
    """Sequential model of cc_init_kernel + cc_merge26_kernel (postproc_kernels.cu): 32-voxel segments linked without
    unions, then one union per (run, neighbouring run) pair issued by the rule in the kernel's comment.  Returns the root
    (smallest voxel index of the component) per voxel, -1 for background."""
    H, W, D = fg.shape
    V = H * W * D
    f = fg.ravel()
    L = np.full(V, -1, dtype=np.int64)
    for base in range(0, V, 32):
        for lane in range(min(32, V - base)):
            v = base + lane
            if not f[v]:
                continue
            s = lane
            while s > 0 and f[base + s - 1] and (base + s) % D != 0:
                s -= 1
            L[v] = base + s

    def find(x):
        while L[x] != x:
            x = L[x]
        return x

    def union(a, b):
        a, b = find(a), find(b)
        if a != b:
            L[max(a, b)] = min(a, b)

    for v in range(V):
        if L[v] < 0:
            continue
        d, t = v % D, v // D
        w, h = t % W, t // W
        prev = d > 0 and f[v - 1]
        if prev and v % 32 == 0:
            union(v, v - 1)
        for dh, dw in ((-1, -1), (-1, 0), (-1, 1), (0, -1)):
            hh, ww = h + dh, w + dw
            if hh < 0 or ww < 0 or ww >= W:
                continue
            row = (hh * W + ww) * D
            um = d > 0 and f[row + d - 1]
            u0 = f[row + d]
            up = d + 1 < D and f[row + d + 1]
            if not prev:
                if um:
                    union(v, row + d - 1)
                elif u0:
                    union(v, row + d)
            if up and not u0:
                union(v, row + d + 1)
    return np.array([find(v) if L[v] >= 0 else -1 for v in range(V)]).reshape(fg.shape)


@pytest.mark.parametrize("shape", [(7, 6, 5), (5, 4, 40), (6, 5, 33), (3, 3, 100)])
def test_run_based_merge_rule_gives_the_26_connected_components(shape):
    """The union rule of cc_merge26_kernel (at most one union per pair of adjacent runs) yields exactly the
    26-connected components, for runs that cross the 32-voxel segments and rows that are not multiples of 32."""
    from scipy import ndimage
    g = np.random.default_rng(sum(shape))
    for density in (0.05, 0.3, 0.6, 0.9):
        fg = g.random(shape) < density
        roots = _run_based_components(fg)
        labels, n = ndimage.label(fg, structure=np.ones((3, 3, 3)))
        pairs = set(zip(labels[fg].tolist(), roots[fg].tolist()))
        assert len(pairs) == n == len(set(roots[fg].tolist()))               # same partition
        for lab, root in pairs:                                             # root = first voxel in raster order
            assert root == int(np.flatnonzero(labels.ravel() == lab)[0])
