"""CPU: host logic of the sharded sliding-window driver (tiling plan, window sharding, the
vote all-reduce under gloo with world_size 2) against the restated MONAI algorithm in
oracle/sliding_window.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lintransunet_b200.sliding_window import reduce_votes, scan_plan, shard_windows
from oracle import sliding_window as OSW


def test_plan_matches_restated_monai_and_config5():
    padded, pad_before, roi, starts = scan_plan((512, 512, 256), (128, 128, 128), 0.5)
    assert padded == (512, 512, 256) and pad_before == (0, 0, 0) and len(starts) == 147      # 7 x 7 x 3
    ref = OSW.dense_patch_starts((512, 512, 256), (128, 128, 128), OSW.get_scan_interval((512, 512, 256), (128,) * 3, 0.5))
    assert starts == ref
    # reference inference geometry: full-plane windows, 1-D tiling over depth, interval int(32*0.4)=12
    _, _, _, s2 = scan_plan((512, 512, 100), (512, 512, 32), 0.6)
    assert [s[2] for s in s2] == [0, 12, 24, 36, 48, 60, 68] and all(s[:2] == (0, 0) for s in s2)
    for img, r, ov in [((70, 64, 33), (32, 32, 16), 0.25), ((16, 40, 8), (32, 32, 16), 0.5), ((96, 96, 96), (96, 96, 96), 0.6)]:
        padded, pb, roi, st = scan_plan(img, r, ov)
        ref = OSW.dense_patch_starts(padded, roi, OSW.get_scan_interval(padded, roi, ov))
        assert st == ref and all(0 <= s[i] <= padded[i] - roi[i] for s in st for i in range(3))


def test_shards_partition_the_windows():
    for n, world in [(147, 1), (147, 2), (147, 4), (147, 8), (5, 8)]:
        parts = [shard_windows(n, r, world) for r in range(world)]
        assert sorted(i for p in parts for i in p) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _fake_predictor(x):
    """Deterministic one-hot 'model': class = bucket of the window's local mean intensity pattern."""
    score = torch.stack([x[:, 0], -x[:, 0], 0.3 * torch.ones_like(x[:, 0])], 1)
    idx = score.argmax(1, keepdim=True)
    return torch.zeros_like(score).scatter_(1, idx, 1.0)


def _votes_for(vol, starts, roi, idxs, C):
    votes = torch.zeros((C,) + tuple(vol.shape), dtype=torch.uint8)
    for i in idxs:
        h, w, d = starts[i]
        win = vol[h:h + roi[0], w:w + roi[1], d:d + roi[2]][None, None]
        lab = _fake_predictor(win)[0].argmax(0)
        for c in range(C):
            votes[c, h:h + roi[0], w:w + roi[1], d:d + roi[2]] += (lab == c).to(torch.uint8)
    return votes


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    vol = torch.randn(40, 36, 24, generator=torch.Generator().manual_seed(3))
    _, _, roi, starts = scan_plan(vol.shape, (16, 16, 16), 0.5)
    votes = _votes_for(vol, starts, roi, shard_windows(len(starts), rank, world), 3)
    reduce_votes(votes)                       # the path's single collective
    if rank == 0:
        q.put(votes.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_vote_reduce_equals_single_rank_and_oracle():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    vol = torch.randn(40, 36, 24, generator=torch.Generator().manual_seed(3))
    _, _, roi, starts = scan_plan(vol.shape, (16, 16, 16), 0.5)
    single = _votes_for(vol, starts, roi, range(len(starts)), 3).numpy()
    assert np.array_equal(got, single)                         # bit-exact: integer votes
    frac = single.astype(np.float32) / single.sum(0, keepdims=True).astype(np.float32)
    ref = OSW.sliding_window_inference(vol[None, None], (16, 16, 16), 4, _fake_predictor, overlap=0.5)[0].numpy()
    assert np.allclose(frac, ref, atol=1e-6)


def test_contiguous_sharding_partitions_the_windows():
    from lintransunet_b200.sliding_window import shard_windows_contiguous
    for n, world in [(147, 1), (147, 2), (147, 4), (147, 8), (5, 8), (0, 3)]:
        parts = [shard_windows_contiguous(n, r, world) for r in range(world)]
        assert sorted(i for p in parts for i in p) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
        assert all(p == list(range(p[0], p[0] + len(p))) for p in parts if p)
