"""GPU, world_size 2: the window-sharded sliding-window inference must be BIT-IDENTICAL to the single-rank result
(integer votes; a window's labels do not depend on the batch it ran in).

With >= 2 GPUs the two ranks use NCCL on separate devices, which also covers the label-only exchange (reduce-scatter of
the votes by H-slab + all-gather of the label slabs).  On a single-GPU box both ranks share cuda:0 and talk through
gloo (NCCL refuses two ranks on one device): that covers the sharding, the balanced batches and the vote all-reduce.
tools/mgpu_check.py is the same check as a torchrun script for 4 and 8 GPUs; bench.py repeats it on every N > 1 run."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, backend, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    local = rank if backend == "nccl" else 0
    torch.cuda.set_device(local)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    from lintransunet_b200 import MaskTransUnet
    from lintransunet_b200.sliding_window import sliding_window_inference
    torch.manual_seed(0)
    m = MaskTransUnet([16, 32, 64, 128, 256], [100, 65, 40, 25, 10], [False, True, True, True, True], 1, 3).cuda().eval()
    vol_host = torch.randn(1, 1, 128, 96, 48, generator=torch.Generator().manual_seed(1)).pin_memory()
    vol = vol_host.cuda()
    roi = (64, 64, 32)                                               # 3 x 2 x 2 = 12 windows, 6 per rank: batches 3+3
    ok = True
    for prec in ("bf16", "fp32"):
        m.precision = prec
        frac_s, lab_s = sliding_window_inference(vol, roi, 4, m, overlap=0.5, return_labels=True, distributed=False)
        frac_d, lab_d = sliding_window_inference(vol, roi, 4, m, overlap=0.5, return_labels=True)
        lab_o = sliding_window_inference(vol, roi, 4, m, overlap=0.5, labels_only=True)
        lab_h = sliding_window_inference(vol_host, roi, 4, m, overlap=0.5, labels_only=True)
        slab, off = sliding_window_inference(vol, roi, 4, m, overlap=0.5, labels_only=True, gather_labels=False)
        ok &= torch.equal(frac_d, frac_s) and torch.equal(lab_d, lab_s) and torch.equal(lab_o, lab_s)
        ok &= torch.equal(lab_h, lab_s)
        ok &= torch.equal(slab, lab_s[:, off:off + slab.shape[1]])
    flag = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        q.put(bool(flag.item()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one_rank_bit_exact():
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, backend, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        got = q.get(timeout=600)
    finally:
        for p in procs:
            p.join(timeout=120)
            if p.is_alive():
                p.kill()
    assert all(p.exitcode == 0 for p in procs)
    assert got is True
    print(f"\n[2 ranks, {backend}] sharded sliding window == single rank, bit-exact (bf16 and fp32)")
