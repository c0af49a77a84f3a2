"""Patch-sharded sliding-window inference over a full CT volume (BASELINE config 5).

Drop-in for the way the reference calls MONAI 0.7.0
``sliding_window_inference(images, roi, sw_batch_size, model, overlap=..., sigma_scale=0)``
(inference_multi_classes.py:143, inference_embed_attn.py:141, utils/utils_3D_embed_full.py:148)
with the default ``mode="constant"``: same tiling, same stitched result
``sum_w pred_w / count`` -- but

* the predictor is the B200 ``MaskTransUnet`` in eval mode, whose output is a one-hot argmax
  (model/trans_3DUnet.py:196-202), so the stitched volume is a vote fraction k/n and is
  accumulated EXACTLY as uint8 vote counts on the device (``ltu_vote_accumulate``);
* windows are independent units: with ``torch.distributed`` initialised they are dealt
  round-robin to the ranks (one process per GPU, weights resident) and the only collective is
  one NCCL sum all-reduce of the uint8 vote volume over NVLink.  Integer votes make the N-GPU
  result bit-identical to the 1-GPU result.

MONAI is not installed in this image; its tiling rules are restated in ``scan_plan`` (the same
restatement lives in oracle/sliding_window.py, which is the checker -- "parity unpinned").
"""
from __future__ import annotations

import math
import os
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import ops
from .unet import MaskTransUnet

SW_STREAMS = max(1, int(os.environ.get("LTU_SW_STREAMS", "3")))     # forwards in flight per rank (A/B switch: 1 = one stream)
_STREAMS = {}


def _side_streams(dev, n):
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), n)
    if key not in _STREAMS:
        _STREAMS[key] = [torch.cuda.Stream(device=dev) for _ in range(n)]
    return _STREAMS[key]


__all__ = ["scan_plan", "shard_windows", "reduce_votes", "max_coverage", "balanced_batches", "sliding_window_inference"]


def scan_plan(image_size: Sequence[int], roi_size: Sequence[int], overlap: float):
    """MONAI 0.7.0 inferers/utils.py::_get_scan_interval + data/utils.py::dense_patch_slices.

    Returns (padded_size, pad_before, starts) where starts is the list of window origins in C order
    of (H, W, D)."""
    if not 0 <= overlap < 1:
        raise ValueError("overlap must be in [0, 1)")
    nd = len(image_size)
    roi = tuple(int(r) if r and r > 0 else int(image_size[i]) for i, r in enumerate(roi_size))   # fall_back_tuple
    padded = tuple(max(int(image_size[i]), roi[i]) for i in range(nd))
    pad_before = tuple((padded[i] - int(image_size[i])) // 2 for i in range(nd))
    interval = []
    for i in range(nd):
        if roi[i] == padded[i]:
            interval.append(roi[i])
        else:
            iv = int(roi[i] * (1 - overlap))
            interval.append(iv if iv > 0 else 1)
    per_dim: List[List[int]] = []
    for i in range(nd):
        num = int(math.ceil(float(padded[i]) / interval[i]))
        scan_dim = next((d for d in range(num) if d * interval[i] + roi[i] >= padded[i]), None)
        n = scan_dim + 1 if scan_dim is not None else 1
        s = []
        for idx in range(n):
            st = idx * interval[i]
            st -= max(st + roi[i] - padded[i], 0)
            s.append(st)
        per_dim.append(s)
    starts = [(h, w, d) for h in per_dim[0] for w in per_dim[1] for d in per_dim[2]]
    return padded, pad_before, roi, starts


def shard_windows(n_windows: int, rank: int, world: int) -> List[int]:
    """Round-robin deal of window indices: rank r owns r, r+world, ..."""
    return list(range(rank, n_windows, world))


def shard_windows_contiguous(n_windows: int, rank: int, world: int) -> List[int]:
    """Contiguous blocks in C order of (H, W, D): rank r owns a slab of the volume along H, so a rank that streams
    the volume from host memory uploads only the rows its windows touch.  Block sizes differ by at most one."""
    base, extra = divmod(n_windows, world)
    lo = rank * base + min(rank, extra)
    return list(range(lo, lo + base + (1 if rank < extra else 0)))


def max_coverage(padded: Sequence[int], roi: Sequence[int], starts) -> int:
    """Largest number of windows covering one voxel = product over the axes of the largest number of distinct window
    origins covering one coordinate.  The votes are uint8 (and packed four to a 32-bit atomicAdd), so it must stay
    <= 255; MONAI's signature accepts any overlap < 1 (e.g. 0.86 in 3-D gives 8^3 = 512 covering windows)."""
    cov = 1
    for ax in range(len(padded)):
        origins = sorted({s[ax] for s in starts})
        best, lo = 1, 0
        for hi, o in enumerate(origins):            # origins within roi of each other cover a common coordinate
            while origins[lo] + roi[ax] <= o:
                lo += 1
            best = max(best, hi - lo + 1)
        cov *= best
    return cov


def balanced_batches(n: int, sw_batch_size: int) -> List[int]:
    """Sizes of the forward batches for n windows: the fewest batches of at most sw_batch_size windows, sized as
    equally as possible (19 windows at sw_batch_size 8 run as 7+6+6, not 8+8+3: a batch of 3 leaves most kernels'
    grids below one wave of the 148 SMs)."""
    if n <= 0:
        return []
    k = -(-n // sw_batch_size)
    base, extra = divmod(n, k)
    return [base + 1] * extra + [base] * (k - extra)


last_h2d_bytes = 0      # bytes uploaded by this process in the most recent call with a host-memory input (bench.py e2e)


def reduce_votes(votes: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the per-rank uint8 vote volumes (the path's single exchange step)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(votes, op=dist.ReduceOp.SUM, group=group)
    return votes


@torch.no_grad()
def sliding_window_inference(inputs: torch.Tensor, roi_size: Sequence[int], sw_batch_size: int,
                             predictor: MaskTransUnet, overlap: float = 0.25, mode: str = "constant",
                             sigma_scale: float = 0.125, padding_mode: str = "constant", cval: float = 0.0,
                             sw_device=None, device=None, *, group=None, distributed: Optional[bool] = None,
                             return_labels: bool = False, return_votes: bool = False, labels_only: bool = False,
                             gather_labels: bool = True):
    """Same positional signature as monai.inferers.sliding_window_inference (0.7.0).

    inputs: fp32 [B, 1, H, W, D] on the GPU, or in (pinned) HOST memory: the volume is then streamed to the device in
    slabs along H on a copy stream while the first windows are already being computed, and under torch.distributed
    each rank owns a contiguous block of windows and uploads only the rows they touch.  Every computation runs on the
    GPU either way.  Returns fp32 [B, C, H, W, D] vote fractions (what MONAI's ``output_image / count_map`` yields for a
    one-hot predictor) or, with ``return_labels``, the uint8 argmax [B, H, W, D] as well.  ``return_votes`` returns the
    exact uint8 vote counts [B, C, H, W, D] INSTEAD of the fractions (the decisions of the inference scripts --
    threshold, round, argmax -- are taken on them directly, see lintransunet_b200/inference.py).

    ``labels_only`` returns just the uint8 argmax [B, H, W, D] and never materialises the fp32 fractions; under
    torch.distributed (NCCL) the exchange is then a reduce-scatter of the votes by H-slab, the argmax runs on the owned
    slab only and the label slabs are all-gathered (67 MB instead of a 201 MB all-reduce for 3x512x512x256).  With
    ``gather_labels=False`` each rank keeps its own slab: the call returns ``(labels_slab [B, H/world, W, D],
    h_offset)``."""
    global last_h2d_bytes
    if str(mode).lower().endswith("gaussian"):
        raise NotImplementedError("only the constant blend mode used by the reference scripts is implemented")
    if not isinstance(predictor, MaskTransUnet):
        raise TypeError("predictor must be a lintransunet_b200.MaskTransUnet (eval-mode one-hot votes)")
    if inputs.dim() != 5 or inputs.shape[1] != 1:
        raise ValueError("inputs must be [B, 1, H, W, D]")
    host_input = not inputs.is_cuda
    if host_input:
        if not torch.cuda.is_available():
            raise RuntimeError("sliding_window_inference needs a CUDA device (no CPU fallback)")
        dev = torch.device(device) if device is not None else next(predictor.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("sliding_window_inference runs on the GPU only: move the predictor to a CUDA device")
    B = inputs.shape[0]
    image_size = tuple(int(s) for s in inputs.shape[2:])
    padded, pad_before, roi, starts = scan_plan(image_size, roi_size, overlap)
    if host_input and (padded != image_size or inputs.dtype != torch.float32):
        inputs = inputs.to(dev, non_blocking=True)      # small / odd volumes: plain upload, then the device path
        last_h2d_bytes = inputs.numel() * inputs.element_size()
        host_input = False
    if padded != image_size:           # volume smaller than the window: symmetric constant pad (MONAI)
        pad = []
        for k in (2, 1, 0):
            diff = padded[k] - image_size[k]
            pad.extend([diff // 2, diff - diff // 2])
        inputs = torch.nn.functional.pad(inputs, pad, mode=padding_mode, value=cval)
    if not host_input:
        inputs = inputs.contiguous().float()
        dev = inputs.device
    if distributed is None:
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if distributed else (0, 1)
    # host-resident volume on >= 4 ranks: every rank uploads a disjoint 1/world slab once and the slabs are all-gathered
    # over NVLink (contiguous window blocks would upload their overlapping halo rows once per neighbour: 738 MB
    # instead of 268 MB at 8 ranks); below that the volume is streamed under the first windows
    gather_upload = (host_input and distributed and world >= 4 and dist.get_backend(group) == "nccl"
                     and padded[0] % world == 0)
    stream_upload = host_input and not gather_upload
    mine = (shard_windows_contiguous if stream_upload else shard_windows)(len(starts), rank, world)
    C = predictor.dim_output
    cov = max_coverage(padded, roi, starts)
    if cov > 255:
        raise ValueError(f"overlap {overlap}: a voxel is covered by up to {cov} windows, the uint8 vote volume holds 255")
    slab_exchange = (labels_only and distributed and dist.get_backend(group) == "nccl" and padded == image_size
                     and padded[0] % world == 0)
    starts_dev = torch.tensor([starts[i] for i in mine], dtype=torch.int32, device=dev).reshape(-1, 3)
    fracs, labels_out = [], []
    if host_input:
        last_h2d_bytes = 0
    if gather_upload:
        src = inputs.contiguous()
    if stream_upload:
        copy_stream = torch.cuda.Stream(device=dev)
        main_stream = torch.cuda.current_stream(dev)
        slab = roi[0] // 2 if roi[0] >= 2 else 1                      # rows per upload: half a window
        src = inputs.contiguous()
    for b in range(B):
        votes = torch.zeros((C,) + padded, dtype=torch.uint8, device=dev)
        if gather_upload:
            rows = padded[0] // world
            part = torch.empty((rows,) + padded[1:], dtype=torch.float32, device=dev)
            part.copy_(src[b, 0, rank * rows:(rank + 1) * rows], non_blocking=True)
            last_h2d_bytes += part.numel() * 4
            vol = torch.empty(padded, dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(vol, part, group=group)
        elif stream_upload:
            # rows [h_lo, h_hi) this rank's windows touch; uploaded in `slab`-row pieces in window order
            h_lo = min(starts[i][0] for i in mine) if mine else 0
            h_hi = max(starts[i][0] for i in mine) + roi[0] if mine else 0
            vol = torch.empty(padded, dtype=torch.float32, device=dev)
            ready = {}                                                    # last row of a piece -> event
            copy_stream.wait_stream(main_stream)
            with torch.cuda.stream(copy_stream):
                for r0 in range(h_lo, h_hi, slab):
                    r1 = min(r0 + slab, h_hi)
                    vol[r0:r1].copy_(src[b, 0, r0:r1], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                    ready[r1] = ev
                    last_h2d_bytes += (r1 - r0) * padded[1] * padded[2] * 4
            vol.record_stream(copy_stream)
            piece_ends = sorted(ready)
        else:
            vol = inputs[b, 0]
        g0 = 0
        batches = balanced_batches(len(mine), sw_batch_size)
        # Two forwards in flight: consecutive batches alternate between two streams, each replaying its own instance of the
        # forward's CUDA graph (predictor.graph_slot).  A forward spends about a quarter of its time in layers that cannot fill
        # the GPU (the small bridges, the coarse convolutions: 2.9 ms of 12.2 ms at batch 8, tools/batch_sweep.py); those
        # now run under the other batch's large kernels.  The votes are integer atomics: the result does not depend on the
        # interleaving.
        n_streams = SW_STREAMS if (len(batches) > 1 and hasattr(predictor, "graph_slot")
                                   and getattr(predictor, "use_cuda_graphs", False)) else 1
        cur = torch.cuda.current_stream(dev)
        lanes = [cur] if n_streams == 1 else _side_streams(dev, n_streams)
        for s in lanes:
            if s is not cur:
                s.wait_stream(cur)
        for k, nb in enumerate(batches):
            st = starts_dev[g0:g0 + nb].contiguous()
            lane = lanes[k % n_streams]
            with torch.cuda.stream(lane):
                if stream_upload:                                         # wait for the last row this batch reads
                    need = max(starts[i][0] for i in mine[g0:g0 + nb]) + roi[0]
                    lane.wait_event(ready[next(e for e in piece_ends if e >= need)])
                if n_streams > 1:
                    predictor.graph_slot = k % n_streams
                    st.record_stream(lane)
                win = ops.gather_windows(vol, st, roi)
                lab = predictor.predict_labels(win)
                ops.vote_accumulate(lab, st, votes)
            g0 += nb
        if n_streams > 1:
            predictor.graph_slot = 0
            for s in lanes:
                cur.wait_stream(s)
                vol.record_stream(s)
                votes.record_stream(s)
        if slab_exchange:
            # reduce-scatter by H-slab (one collective per class volume: votes[c] is contiguous [H,W,D]), argmax on the
            # owned slab, then (optionally) all-gather the uint8 label slabs
            rows = padded[0] // world
            mine_v = torch.empty((C, rows) + padded[1:], dtype=torch.uint8, device=dev)
            for c in range(C):
                dist.reduce_scatter_tensor(mine_v[c], votes[c], op=dist.ReduceOp.SUM, group=group)
            lab_slab = ops.vote_argmax(mine_v)
            if gather_labels:
                full = torch.empty(padded, dtype=torch.uint8, device=dev)
                dist.all_gather_into_tensor(full, lab_slab, group=group)
                labels_out.append(full)
            else:
                labels_out.append(lab_slab)
            continue
        if distributed:
            reduce_votes(votes, group)
        if padded != image_size:
            sl = tuple(slice(pad_before[i], pad_before[i] + image_size[i]) for i in range(3))
            votes = votes[(slice(None),) + sl].contiguous()
        if not labels_only:
            fracs.append(votes if return_votes else ops.vote_fractions(votes))
        if return_labels or labels_only:
            labels_out.append(ops.vote_argmax(votes))
    if labels_only:
        lab = torch.stack(labels_out, 0) if B > 1 else labels_out[0].unsqueeze(0)
        if slab_exchange and not gather_labels:
            return lab, rank * (padded[0] // world)
        return (lab, 0) if not gather_labels else lab
    out = torch.stack(fracs, 0) if B > 1 else fracs[0].unsqueeze(0)
    if return_labels:
        return out, (torch.stack(labels_out, 0) if B > 1 else labels_out[0].unsqueeze(0))
    return out
