"""Forward + backward device time of the two model slices whose backward exists, at BASELINE config 2's
shape (batch 2 of 96^3 patches, bf16): the Encoder and bridge 1's EmbedAttention3DBlock.  CUDA events, eager launches.

    python tools/train_slice_probe.py > gpurun_out/train_slice_probe.md"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lintransunet_b200 import _native, ops  # noqa: E402
from lintransunet_b200.backward import (embed_block_backward, embed_block_train, encoder_backward,  # noqa: E402
                                        encoder_train)
from lintransunet_b200.unet import EmbedAttention3DBlock, Encoder  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def main():
    torch.manual_seed(0)
    print("| slice | shape | forward (training mode) ms | backward ms | native launches fwd + bwd |")
    print("|---|---|---:|---:|---:|")
    enc = Encoder([16, 32, 64, 128, 256], 1).cuda()
    x = torch.randn(2, 1, 96, 96, 96, device="cuda")
    n0 = _native.launch_count()
    t_f, (bottle, skips, saved) = timed(lambda: encoder_train(x, enc))
    d_b = torch.randn_like(bottle)
    d_s = [torch.randn_like(s) for s in skips]
    t_b, _ = timed(lambda: encoder_backward(d_b, d_s, saved))
    n = (_native.launch_count() - n0) // 4
    print(f"| Encoder (stem + 4 DownBlocks) | 2 x 1 x 96^3 | {t_f:.2f} | {t_b:.2f} | {n} |", flush=True)
    blk = EmbedAttention3DBlock(32, 128, 4, 8).cuda()
    xb = torch.randn(2, 78, 46, 96, 32, device="cuda").to(torch.bfloat16)
    n0 = _native.launch_count()
    t_f, (y, saved) = timed(lambda: embed_block_train(xb, blk))
    dy = torch.randn_like(y)
    t_b, _ = timed(lambda: embed_block_backward(dy, saved))
    n = (_native.launch_count() - n0) // 4
    print(f"| EmbedAttention3DBlock, bridge 1 (43 056 tokens / sample, d_model 128) | 2 x 78x46x96 x 32 | {t_f:.2f} | {t_b:.2f} | {n} |",
          flush=True)
    print("\nEager launches with per-call weight repacking (no plan cache, no CUDA graph yet): an upper bound of the device time.")


if __name__ == "__main__":
    main()
