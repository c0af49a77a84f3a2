"""Sharded inference driver: what inference_embed_attn.py (binary) and inference_multi_classes.py
(multi-class) of the reference do per patient, without MONAI and without leaving the GPU
(the connected-component step is 8f-4).

Per case (reference lines in brackets):

1. CT volume ``image/<name>.npy`` stored [D,H,W]: clip HU, z-score, -> [1,1,H,W,D]
   [dataset/CT_pancreas_ids.py:219-251, dataset/CT_pancreas_multi_class.py:222-254];
2. window-sharded sliding window over the ranks -> exact uint8 votes [C,H,W,D]
   [inference_*.py:141/:143 ``sliding_window_inference(images, (roi, roi, depth), sw_batch_size,
   model, overlap=0.6, sigma_scale=0)``];
3. binary: ``(predict >= 0.5)`` [inference_embed_attn.py:147];
   multi-class: ``torch.round(predict)`` -> ``KeepLargestConnectedComponent(applied_labels=[1,2],
   independent=False, connectivity=3)`` -> ``predict2[:,0] = 1 - predict2[:,1] - predict2[:,2]``
   [inference_multi_classes.py:104,:148-152];
4. metrics [criterion_list defaults, inference_embed_attn.py:62-64 / inference_multi_classes.py:57-59]
   from one integer-count kernel;
5. ``np.save`` of the class-1 mask (binary, float32) or the argmax (multi-class, int64), permuted back to
   [D,H,W] [inference_embed_attn.py:152-158, inference_multi_classes.py:156-162].

Run it like the reference scripts, or under torchrun for one process per GPU:

    python -m lintransunet_b200.inference --dir_data DATA --pretrained_dir model.pt --dim_output 3 --is_save
    python -m torch.distributed.run --nproc-per-node 8 -m lintransunet_b200.inference ...

Steps 2-4 run through libltu_b200.so; there is no CPU path.
"""
from __future__ import annotations

import argparse
import json
import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from . import ops
from .sliding_window import sliding_window_inference
from .unet import MaskTransUnet, get_model_dict

__all__ = ["prepare_ct", "metrics_from_counts", "postprocess_votes", "segment_volume", "prediction_array", "main"]

# dataset/CT_pancreas_ids.py:193-197 (binary) and dataset/CT_pancreas_multi_class.py:197-200 (multi-class)
CT_NORM = {False: dict(low_clip=-91.0, high_clip=250.0, mean=86.9, std=39.4),
           True: dict(low_clip=-96.0, high_clip=215.0, mean=77.99, std=75.4)}


def prepare_ct(volume_dhw: np.ndarray, multi_class: bool) -> torch.Tensor:
    """[D,H,W] HU -> fp32 [1,1,H,W,D] pinned host tensor, clipped and z-scored like EvaPanCTDataset.__getitem__."""
    n = CT_NORM[bool(multi_class)]
    img = np.array(volume_dhw, copy=True)                 # same statements, same numpy promotion as the reference
    img[img < n["low_clip"]] = n["low_clip"]
    img[img > n["high_clip"]] = n["high_clip"]
    img = (img - n["mean"]) / n["std"]
    img = img.astype(np.float32)
    t = torch.from_numpy(img).permute(1, 2, 0).contiguous()[None, None]
    return t.pin_memory() if torch.cuda.is_available() else t


def _profile_distance(p_rows: torch.Tensor, t_rows: torch.Tensor, eps: float, scale: float, thr: float = 10.0) -> float:
    """LocalizationLoss on the H-profiles (loss/criterions.py:192-241: the three loop iterations are the same H
    profile; loss/multi_criterions.py:232-281)."""
    p = torch.sigmoid(p_rows.float() - thr)
    t = torch.sigmoid(t_rows.float() - thr)
    dp = torch.cumsum(p, -1) / (p.sum(-1, keepdim=True) + eps)
    dt = torch.cumsum(t, -1) / (t.sum(-1, keepdim=True) + eps)
    return float(scale * torch.mean(torch.abs(dp - dt)))


def metrics_from_counts(counts: torch.Tensor, multi_class: bool) -> Dict[str, float]:
    """counts int64 [C+1,H,3] (ops.overlap_counts) -> the values the reference prints, keyed by its criterion names.
    Sums of 0/1 voxels are exact integers here; the reference adds them in fp32, which is exact below 2^24 voxels
    per class, and the final ratios are formed in fp32 like the reference's."""
    c = counts.detach().cpu()
    C = c.shape[0] - 1
    tot = c.sum(1).to(torch.float32)                       # [C+1, 3] = TP, P, T

    def dice(row, eps=1e-9):
        tp, p, t = tot[row]
        return float(1 - (2 * tp + eps) / (p + t + eps))

    def rec(row, eps=1e-5):
        tp, _, t = tot[row]
        return float((tp + eps) / (t + eps))

    def prec(row, eps=1e-5):
        tp, p, _ = tot[row]
        return float((tp + eps) / (p + eps))

    out: Dict[str, float] = {}
    if not multi_class:
        out["DiceClassLoss"], out["Recall"], out["Precision"] = dice(1), rec(1), prec(1)
        out["LocalizationLoss"] = _profile_distance(c[1, :, 1], c[1, :, 2], 1e-6, 8.0)
        return out
    out["DiceClassLoss0"] = dice(C)                        # foreground = 1 - channel 0
    for k in range(1, C):
        sfx = "" if k == 1 else str(k)
        out[f"DiceClassLoss{sfx}"] = dice(k)
    for k in range(1, C):
        sfx = "" if k == 1 else str(k)
        out[f"Recall{sfx}"], out[f"Precision{sfx}"] = rec(k), prec(k)
    out["LocalizationLoss"] = _profile_distance(c[C, :, 1], c[C, :, 2], 1e-6, 1.0)
    return out


def postprocess_votes(votes: torch.Tensor, multi_class: bool, threshold: float = 0.5,
                      applied_labels: Optional[Sequence[int]] = None, connectivity: int = 3,
                      keep_largest: bool = True) -> torch.Tensor:
    """uint8 votes [C,H,W,D] -> uint8 one-hot [C,H,W,D] = the reference's ``predict2`` (step 3 of the module doc)."""
    C = votes.shape[0]
    if not multi_class:
        return ops.vote_decide(votes, ops.DECIDE_THRESHOLD, threshold)
    onehot = ops.vote_decide(votes, ops.DECIDE_ROUND)
    if keep_largest:
        ops.keep_largest_component_(onehot, list(applied_labels) if applied_labels is not None else list(range(1, C)),
                                    connectivity=connectivity, independent=False)
    onehot[0] = 1 - onehot[1:].sum(0, dtype=torch.uint8)    # at most one class rounds to 1: stays in {0, 1}
    return onehot


@torch.no_grad()
def segment_volume(model: MaskTransUnet, image: torch.Tensor, roi_size: Sequence[int], sw_batch_size: int = 4,
                   overlap: float = 0.6, multi_class: Optional[bool] = None, threshold: float = 0.5,
                   keep_largest: bool = True, connectivity: int = 3, group=None) -> torch.Tensor:
    """image fp32 [1,1,H,W,D] (GPU, or pinned host memory) -> uint8 one-hot [C,H,W,D] on the GPU.  Under
    torch.distributed every rank must call it: the windows are sharded, every rank gets the full result."""
    if image.dim() != 5 or image.shape[0] != 1:
        raise ValueError("segment_volume takes one case [1,1,H,W,D] at a time (the reference uses batch_size 1)")
    multi = model.dim_output > 2 if multi_class is None else bool(multi_class)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        votes = sliding_window_inference(image, roi_size, sw_batch_size, model, overlap=overlap, sigma_scale=0,
                                         group=group, return_votes=True)[0]
    return postprocess_votes(votes, multi, threshold, connectivity=connectivity, keep_largest=keep_largest)


def prediction_array(onehot: torch.Tensor, multi_class: bool) -> np.ndarray:
    """What the reference saves: class-1 mask float32 (inference_embed_attn.py:152-158) or argmax int64
    (inference_multi_classes.py:156-162), permuted (2,0,1) back to [D,H,W]."""
    if multi_class:
        out = ops.vote_argmax(onehot.contiguous())           # first maximum wins, like torch.argmax(predict2, dim=1)
        return out.permute(2, 0, 1).cpu().numpy().astype(np.int64)
    return onehot[1].float().permute(2, 0, 1).cpu().numpy()


def evaluate_case(onehot: torch.Tensor, label_dhw: np.ndarray, multi_class: bool) -> Dict[str, float]:
    """label [D,H,W] as stored on disk -> the reference's per-patient metric values."""
    lab = np.asarray(label_dhw)
    lab = lab.astype(np.int64) if multi_class else (lab > 0.5)
    target = torch.from_numpy(np.ascontiguousarray(lab.astype(np.uint8))).permute(1, 2, 0).contiguous().to(onehot.device)
    return metrics_from_counts(ops.overlap_counts(onehot.contiguous(), target), multi_class)


def _int_list(text):
    return [int(v) for v in str(text).replace("[", "").replace("]", "").split(",") if v.strip()]


def _bool_list(text):
    return [v.strip().lower() in ("1", "true", "t", "yes") for v in str(text).replace("[", "").replace("]", "").split(",")
            if v.strip()]


def get_parse(argv=None):
    """Argument names of the reference scripts (inference_multi_classes.py:18-67); list arguments are comma separated."""
    p = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    p.add_argument("--dir_data", type=str, required=True, help="folder with image/*.npy and (optionally) label/*.npy")
    p.add_argument("--pretrained_dir", type=str, default=None,
                   help="state_dict file, or the reference's log folder holding fold_<k>/temp_model.pt")
    p.add_argument("--model_name", type=str, default="MaskTransUnet")
    p.add_argument("--depth_size", type=int, default=32)
    p.add_argument("--num_layers", type=_int_list, default=[16, 32, 64, 128, 256])
    p.add_argument("--roi_size_list", type=_int_list, default=[100, 65, 40, 25, 10])
    p.add_argument("--is_roi_list", type=_bool_list, default=[False, True, True, True, True])
    p.add_argument("--dim_input", type=int, default=1)
    p.add_argument("--dim_output", type=int, default=2)
    p.add_argument("--kernel_size", type=int, default=3)
    p.add_argument("--device", type=str, default="cuda")
    p.add_argument("--is_save", action="store_true")
    p.add_argument("--saved_folder", type=str, default="./prediction/test")
    # constants the reference hard-codes in main() (inference_*.py:96-101)
    p.add_argument("--roi_size", type=int, default=512)
    p.add_argument("--sw_batch_size", type=int, default=4)
    p.add_argument("--overlap", type=float, default=0.6)
    p.add_argument("--threshold", type=float, default=0.5)
    p.add_argument("--split_json", type=str, default=None, help="split_dataset_8.json of the reference (optional)")
    p.add_argument("--fold", type=int, default=0)
    p.add_argument("--summary_json", type=str, default="summary_4_fold.json")
    p.add_argument("--seed", type=int, default=0, help="random-init seed when no --pretrained_dir is given")
    return p.parse_args(argv)


def _load_model(args, device: torch.device) -> MaskTransUnet:
    model_fn = get_model_dict(args.model_name)
    model = model_fn(num_layers=args.num_layers, roi_size_list=args.roi_size_list, is_roi_list=args.is_roi_list,
                     dim_input=args.dim_input, dim_output=args.dim_output, kernel_size=args.kernel_size)
    if args.pretrained_dir:
        path = args.pretrained_dir
        if os.path.isdir(path):
            path = os.path.join(path, f"fold_{args.fold}", "temp_model.pt")
        state = torch.load(path, map_location="cpu")
        if not isinstance(state, dict):                     # a pickled module (the commented-out variant, :85)
            state = state.state_dict()
        state = {k[7:] if k.startswith("module.") else k: v for k, v in state.items()}    # saved through DataParallel
        model.load_state_dict(state)
    return model.to(device).eval()


def main(argv=None) -> Dict[str, object]:
    args = get_parse(argv)
    if not torch.cuda.is_available():
        raise RuntimeError("lintransunet_b200.inference needs a CUDA device (sm_100a); there is no CPU path")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=device)
    multi = args.dim_output > 2
    torch.manual_seed(args.seed)
    model = _load_model(args, device)

    names = sorted(os.listdir(os.path.join(args.dir_data, "image")))
    label_dir = os.path.join(args.dir_data, "label")
    labels = sorted(os.listdir(label_dir)) if os.path.isdir(label_dir) else None
    ids = list(range(len(names)))
    if args.split_json:
        with open(args.split_json) as f:
            ids = json.load(f)[f"test_id fold_{args.fold}"][:-1]        # inference_*.py:110-113
    if args.is_save and rank == 0:
        os.makedirs(args.saved_folder, exist_ok=True)

    per_patient: List[Dict[str, float]] = []
    for i in ids:
        name = names[i]
        image = prepare_ct(np.load(os.path.join(args.dir_data, "image", name)), multi)
        onehot = segment_volume(model, image, (args.roi_size, args.roi_size, args.depth_size), args.sw_batch_size,
                                args.overlap, multi_class=multi, threshold=args.threshold)
        if rank != 0:
            continue
        row: Dict[str, float] = {}
        if labels is not None:
            row = evaluate_case(onehot, np.load(os.path.join(label_dir, labels[i])), multi)
            for k, v in row.items():
                print(f"eval patient average {k}", v)
        per_patient.append(row)
        if args.is_save:
            stem = "{:0>4}".format(name) + ("_multi" if multi else "")
            np.save(os.path.join(args.saved_folder, stem), prediction_array(onehot, multi))
    summary: Dict[str, object] = {}
    if rank == 0:
        keys = list(per_patient[0].keys()) if per_patient and per_patient[0] else []
        mean = {k: float(np.mean([r[k] for r in per_patient])) for k in keys}
        for k, v in mean.items():
            print(f"eval total average {k} loss", v)
        summary = {f"patient_{args.fold}": [[r[k] for k in keys] for r in per_patient],
                   f"summary_{args.fold}": [mean[k] for k in keys], "criterions": keys}
        if args.summary_json:
            with open(args.summary_json, "w") as f:
                json.dump(summary, f, indent=4)
    if world > 1:
        dist.barrier()
    return summary


if __name__ == "__main__":
    main()
