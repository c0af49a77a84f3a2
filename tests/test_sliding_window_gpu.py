"""GPU: the sharded sliding-window driver (gather -> forward -> uint8 votes -> fractions) against
the restated MONAI algorithm (oracle/sliding_window.py) driving the CPU oracle model."""
import numpy as np
import pytest
import torch

from oracle import ltu_oracle as O
from oracle import sliding_window as OSW

pytestmark = pytest.mark.gpu


def test_sliding_window_matches_restated_monai_with_oracle_model():
    from lintransunet_b200 import MaskTransUnet
    from lintransunet_b200.sliding_window import sliding_window_inference
    cfg = O.UnetConfig(dim_output=3)
    sd = O.make_state_dict(cfg, seed=2)
    m = MaskTransUnet(list(cfg.num_layers), list(cfg.roi_size_list), list(cfg.is_roi_list), 1, 3, dropout=0.0)
    m.load_state_dict(sd)
    m.cuda().eval()
    m.precision = "fp32"
    vol = O.make_input((1, 1, 96, 96, 24), seed=4, blob=True)
    roi, ov = (64, 64, 16), 0.5                                     # 2 x 2 x 2 = 8 windows, ragged batches of 3
    frac, labels = sliding_window_inference(vol.cuda(), roi, 3, m, overlap=ov, return_labels=True)
    ref = OSW.sliding_window_inference(vol, roi, 3, lambda x: O.mask_trans_unet_forward(x, sd, cfg)["onehot"], overlap=ov)
    assert frac.shape == ref.shape == (1, 3, 96, 96, 24)
    mism = float((frac.cpu() != ref).any(1).float().mean())
    print(f"\n[sliding window fp32] voxels with a differing vote fraction: {mism:.3e}")
    assert mism <= 1e-3                                             # integer votes: equal unless an argmax tie flips
    assert float((frac.sum(1) - 1).abs().max()) < 1e-6
    assert torch.equal(labels[0].long().cpu(), frac[0].argmax(0).cpu())
    # a volume smaller than the window is padded symmetrically and cropped back
    small = O.make_input((1, 1, 64, 32, 16), seed=5)
    f2 = sliding_window_inference(small.cuda(), roi, 2, m, overlap=ov)
    r2 = OSW.sliding_window_inference(small, roi, 2, lambda x: O.mask_trans_unet_forward(x, sd, cfg)["onehot"], overlap=ov)
    assert f2.shape == r2.shape == (1, 3, 64, 32, 16)
    assert float((f2.cpu() != r2).any(1).float().mean()) <= 1e-3


def test_host_memory_input_is_streamed_and_bit_identical():
    """A pinned host volume is uploaded slab by slab on a copy stream; votes and labels equal the device-input result."""
    from lintransunet_b200 import MaskTransUnet, sliding_window as sw
    torch.manual_seed(0)
    m = MaskTransUnet([16, 32, 64, 128, 256], [100, 65, 40, 25, 10], [False, True, True, True, True], 1, 3).cuda().eval()
    vol = torch.randn(1, 1, 160, 96, 32, generator=torch.Generator().manual_seed(2)).pin_memory()
    roi = (64, 64, 16)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        f_dev, l_dev = sw.sliding_window_inference(vol.cuda(), roi, 4, m, overlap=0.5, return_labels=True)
        f_host, l_host = sw.sliding_window_inference(vol, roi, 4, m, overlap=0.5, return_labels=True)
    assert sw.last_h2d_bytes == vol.numel() * 4
    assert torch.equal(f_dev, f_host) and torch.equal(l_dev, l_host)


def test_forwards_in_flight_do_not_change_the_result():
    """Consecutive batches run on alternating streams, each replaying its own instance of the forward's CUDA graph
    (sliding_window.SW_STREAMS): votes are integer atomics and every batch keeps its composition, so labels and vote fractions
    equal the one-stream schedule bit for bit -- device and pinned-host input, repeated calls (graph instances are reused)."""
    from lintransunet_b200 import MaskTransUnet, sliding_window as sw
    torch.manual_seed(0)
    m = MaskTransUnet([16, 32, 64, 128, 256], [100, 65, 40, 25, 10], [False, True, True, True, True], 1, 3).cuda().eval()
    vol = torch.randn(1, 1, 160, 160, 48, generator=torch.Generator().manual_seed(3)).pin_memory()
    roi = (64, 64, 16)                                              # 4 x 4 x 5 = 80 windows: 27 batches of 3
    prev = sw.SW_STREAMS
    out = {}
    try:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            for n in (1, 3, 2):
                sw.SW_STREAMS = n
                for rep in range(2):
                    f, l = sw.sliding_window_inference(vol.cuda(), roi, 3, m, overlap=0.5, return_labels=True)
                    lab_only = sw.sliding_window_inference(vol, roi, 3, m, overlap=0.5, labels_only=True)
                    out[(n, rep)] = (f.clone(), l.clone(), lab_only.clone())
    finally:
        sw.SW_STREAMS = prev
    assert m.graph_slot == 0
    f0, l0, o0 = out[(1, 0)]
    for key, (f, l, o) in out.items():
        assert torch.equal(f, f0) and torch.equal(l, l0) and torch.equal(o, o0), key
    assert torch.equal(l0, o0)
