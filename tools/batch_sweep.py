"""Forward time of the benchmark model against the batch size (CUDA-graph replays, bf16): the per-forward fixed cost that limits
the 8-GPU sliding window (7 + 6 + 6 windows per rank).   python tools/batch_sweep.py"""
import sys
import torch
sys.path.insert(0, ".")
from lintransunet_b200 import MaskTransUnet  # noqa: E402

torch.manual_seed(0)
m = MaskTransUnet([16, 32, 64, 128, 256], [100, 65, 40, 25, 10], [False, True, True, True, True], 1, 3).cuda().eval()
print("| batch | ms per forward | ms per window |\n|---:|---:|---:|")
with torch.autocast("cuda", dtype=torch.bfloat16):
    for B in (1, 2, 3, 4, 6, 7, 8):
        x = torch.randn(B, 1, 128, 128, 128, device="cuda")
        for _ in range(3):
            m.predict_labels(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            m.predict_labels(x)
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 5
        print(f"| {B} | {t:.2f} | {t / B:.2f} |", flush=True)
