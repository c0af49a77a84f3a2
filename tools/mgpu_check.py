#!/usr/bin/env python
"""Multi-GPU check (run under torchrun, one process per GPU): the window-sharded sliding-window
inference with the NCCL uint8 vote all-reduce must be BIT-IDENTICAL to the single-GPU result."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lintransunet_b200 import MaskTransUnet  # noqa: E402
from lintransunet_b200.sliding_window import sliding_window_inference  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.manual_seed(0)
    m = MaskTransUnet([16, 32, 64, 128, 256], [100, 65, 40, 25, 10], [False, True, True, True, True], 1, 3).cuda().eval()
    vol = torch.randn(1, 1, 192, 160, 96, generator=torch.Generator().manual_seed(1)).cuda()
    roi = (64, 64, 32)
    for prec in ("bf16", "fp32"):
        m.precision = prec
        frac_d, lab_d = sliding_window_inference(vol, roi, 4, m, overlap=0.5, return_labels=True)              # sharded
        frac_s, lab_s = sliding_window_inference(vol, roi, 4, m, overlap=0.5, return_labels=True, distributed=False)
        same = torch.equal(frac_d, frac_s) and torch.equal(lab_d, lab_s)
        # label-only exchange (reduce-scatter by H-slab + all-gather of the label slabs) and the per-rank slab
        lab_o = sliding_window_inference(vol, roi, 4, m, overlap=0.5, labels_only=True)
        slab, off = sliding_window_inference(vol, roi, 4, m, overlap=0.5, labels_only=True, gather_labels=False)
        lab_h = sliding_window_inference(vol.cpu().pin_memory(), roi, 4, m, overlap=0.5, labels_only=True)   # host volume
        same = same and torch.equal(lab_o, lab_s) and torch.equal(slab, lab_s[:, off:off + slab.shape[1]])
        same = same and torch.equal(lab_h, lab_s)
        flags = torch.tensor([int(same)], device="cuda")
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"[mgpu world={world} {prec}] sharded == single-GPU bit-exact: {bool(flags.item())}; "
                  f"class histogram {torch.bincount(lab_d.flatten().long(), minlength=3).tolist()}", flush=True)
        assert bool(flags.item()), "sharded result differs from the single-GPU result"
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
