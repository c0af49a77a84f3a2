"""Device times of the post-processing kernels on a config-5 sized volume (3 x 512 x 512 x 256 votes), CUDA events.

    python tools/postproc_probe.py > gpurun_out/postproc_probe.md

The one-hot volume is synthetic: a few large boxes per class plus sparse speckle (what a segmentation looks like to the
component labelling: a handful of big components and many tiny ones)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lintransunet_b200 import ops  # noqa: E402


def timeit(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3)
    return best


def main():
    C, H, W, D = 3, 512, 512, 256
    V = H * W * D
    g = torch.Generator(device="cuda").manual_seed(0)
    lab = torch.zeros(H, W, D, dtype=torch.uint8, device="cuda")
    lab[torch.rand(H, W, D, device="cuda", generator=g) < 0.002] = 1                  # speckle
    lab[180:330, 150:300, 60:200] = 1
    lab[200:260, 290:380, 80:160] = 2
    lab[20:60, 30:70, 10:40] = 2
    n = torch.randint(1, 9, (H, W, D), device="cuda", generator=g, dtype=torch.uint8)  # coverage counts 1..8
    votes = torch.stack([(lab == c).to(torch.uint8) * n for c in range(C)]).contiguous()
    target = torch.roll(lab, shifts=(5, -3, 2), dims=(0, 1, 2)).contiguous()
    onehot = ops.vote_decide(votes, ops.DECIDE_ROUND)

    rows = []
    t = timeit(lambda: ops.vote_decide(votes, ops.DECIDE_ROUND))
    rows.append(("vote_decide (round)", t, 2 * C * V))
    t = timeit(lambda: ops.vote_fractions(votes))
    rows.append(("vote_fractions (fp32 volume, for comparison)", t, 5 * C * V))
    work = onehot.clone()

    def cc():
        work.copy_(onehot)
        ops.keep_largest_component_(work, [1, 2], connectivity=3)
    t_copy = timeit(lambda: work.copy_(onehot))
    t = timeit(cc) - t_copy
    rows.append(("keep_largest_component (5 kernels, 26-connectivity)", t, C * V + 20 * V))
    t = timeit(lambda: ops.overlap_counts(onehot, target))
    rows.append(("overlap_counts", t, (C + 2) * V))
    kept = int(work[1:].sum())
    print(f"# post-processing kernels, votes uint8 [{C},{H},{W},{D}] ({V / 1e6:.0f} M voxels), foreground {int(onehot[1:].sum())} "
          f"voxels -> {kept} kept\n")
    print("| kernel | us | algorithmic MB | GB/s |")
    print("|---|---:|---:|---:|")
    for name, us, nbytes in rows:
        print(f"| {name} | {us:.0f} | {nbytes / 1e6:.0f} | {nbytes / us / 1e3:.0f} |")


if __name__ == "__main__":
    main()
