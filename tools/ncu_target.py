#!/usr/bin/env python
"""Short target for `ncu --set full`: two eager bf16 forwards of the benchmark model
(batch 8 of 128^3 patches = one sliding-window batch), no CUDA graphs."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lintransunet_b200 import MaskTransUnet  # noqa: E402

torch.manual_seed(0)
m = MaskTransUnet([16, 32, 64, 128, 256], [100, 65, 40, 25, 10], [False, True, True, True, True], 1, 3).cuda().eval()
m.use_cuda_graphs = False
x = torch.randn(int(os.environ.get("NCU_BATCH", "8")), 1, 128, 128, 128, device="cuda")
with torch.autocast("cuda", dtype=torch.bfloat16):
    for _ in range(2):
        y = m.predict_labels(x)
torch.cuda.synchronize()
print("ok", y.shape)
