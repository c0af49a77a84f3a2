#!/usr/bin/env python
"""Headline benchmark: BASELINE.json metric "voxels/sec fwd (128^3 patch) at 1/2/4/8 B200;
linear-attn HBM GB/s vs peak".

    python bench.py --gpus N --steps K --warmup W            # this repo (sm_100a kernels)
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference algorithm on the host CPU

A step is one sliding-window inference of the multi-class MaskTransUnet (dim_output=3, random
init, bf16) over a synthetic 512x512x256 CT volume: 128^3 windows, 50 % overlap => 147 windows,
sw_batch_size 8 (BASELINE config 5, which runs the config-4 forward -- batch 8 of 128^3 patches --
19 times, in batches of 8 and 7).  Windows are dealt round-robin to the N ranks (one process per GPU,
torchrun); the only exchange is a reduce-scatter of the uint8 vote volume by H-slab followed by an
all-gather of the uint8 label slabs (NCCL).  Total work is fixed => strong scaling.

  value : window voxels / s = 147 * 128^3 * K / t, volume resident in HBM, CUDA events, max over ranks
  e2e   : same, through lintransunet_b200.sliding_window.sliding_window_inference with the volume in
          pinned host memory (H2D inside the timed region) and every rank's slab of the stitched label
          volume read back to its host (D2H inside the timed region)
  roofline / kernels : per-kernel CUDA-event timings of the timed steps vs MEASURED_PEAKS.json
          (the dominant kernel by time; tensor-bound kernels report algorithmic AND executed TFLOP/s)
  parity : UNTIMED, after the timed region: two windows of the volume through the timed model against
          the oracle (fp32 and its own bf16-autocast floor); at N > 1 the sharded label volume must
          equal rank 0's single-rank result bit for bit; labels_crc32 is the same number at every N
  cpu_baseline : the oracle (CPU port of the reference algorithm) on the box's host cores, 8 of the
          147 windows after one warm-up window (N=1 only, about 10 s)
  gpu_eager_baseline : the same oracle port in eager PyTorch on the same B200 (bf16 autocast and
          fp32, batch 8 x 128^3) -- the kernel-vs-library bar of SURVEY 8d (N=1 only)

The reference arm times the same oracle port (the reference is pure Python/PyTorch and cannot
travel to the GPU box; oracle/ltu_oracle.py is pinned to it by tests/golden) on all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "voxels/sec fwd (128^3 patch)"
VOLUME = (512, 512, 256)
ROI = (128, 128, 128)
OVERLAP = 0.5
SW_BATCH = 8
DIM_OUTPUT = 3
MODEL_CFG = dict(num_layers=[16, 32, 64, 128, 256], roi_size_list=[100, 65, 40, 25, 10],
                 is_roi_list=[False, True, True, True, True], dim_input=1, dim_output=DIM_OUTPUT)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=float(p["hbm_gbs"]), tensor=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                    source="MEASURED_PEAKS.json (hbm_gbs, bf16_tflops_sustained)")
    return dict(hbm=6650.0, tensor=1400.0, source="fallback (B200_PROFILING.md)")


def config_dict(n_gpus):
    return {"workload": "config5: sliding-window inference, synthetic 512x512x256 volume, 128^3 windows, "
                        "50% overlap (147 windows), sw_batch_size 8 (= config4 forward per batch), "
                        "multi-class MaskTransUnet dim_output=3, eval, bf16 autocast",
            "windows": 147, "sw_batch_size": SW_BATCH, "parallelism": f"window-sharded dp{n_gpus}",
            "forwards_in_flight": "3 per rank: consecutive batches of 8 windows alternate between three CUDA streams, each replaying "
                                  "its own instance of the forward's CUDA graph (LTU_SW_STREAMS; the result is bit-identical)",
            "l2": "inputs larger than L2 (268 MB volume, 134 MB+ activations per layer)"}


# --------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU with NVML while the timed region runs."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def run(self):
        while not self.stop_flag and self.nv is not None:
            try:
                self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                bits = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.1)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def visible_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# --------------------------------------------------------------------------- CPU oracle timing
def oracle_window_seconds(steps: int, warmup: int, shape=(1, 1, 128, 128, 128)):
    from oracle import ltu_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.UnetConfig(dim_output=DIM_OUTPUT)
    sd = O.make_state_dict(cfg, seed=0)
    x = O.make_input(shape, seed=1)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.mask_trans_unet_forward(x, sd, cfg)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return times, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times, cores = oracle_window_seconds(args.steps, args.warmup)
    vox = 128 ** 3
    t = sum(times) / len(times)
    sample = "one 1x1x128^3 window forward (fp32) per step, i.e. 1/147 of the volume; oracle port of the reference"
    line = {"impl": "reference", "metric": METRIC, "value": vox / t, "unit": "voxels/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args.gpus),
            "cpu_baseline": {"value": vox / t, "unit": "voxels/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": vox / t, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- our arm
def gpu_eager_baseline(dev, steps=5, warmup=2):
    """The survey's real bar: the reference ALGORITHM in eager PyTorch on the same B200 -- the oracle
    port (device-agnostic torch, pinned to the unmodified reference by tests/golden), one batch of 8 x 128^3 windows
    per step, under bf16 autocast (what the reference scripts do) and in true fp32.  CUDA events, same clocks.
    Checker-side code: it is timed as a BASELINE, never shipped."""
    from oracle import ltu_oracle as O
    cfg = O.UnetConfig(dim_output=DIM_OUTPUT)
    sd = {k: v.to(dev) for k, v in O.make_state_dict(cfg, seed=0).items()}
    x = O.make_input((SW_BATCH, 1) + ROI, seed=1).to(dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    out = {}
    for name, ac in (("bf16_autocast", True), ("fp32", False)):
        ts = []
        with torch.no_grad():
            for i in range(warmup + steps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
                    O.mask_trans_unet_forward(x, sd, cfg)
                e1.record()
                torch.cuda.synchronize()
                if i >= warmup:
                    ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        out[name] = {"ms_per_batch_of_8": round(ms, 3), "voxels_per_s": SW_BATCH * 128 ** 3 / (ms / 1e3)}
    out["what"] = ("oracle port of the reference forward in eager PyTorch (cuDNN / cuBLAS / ATen kernels) on the same "
                   "B200, batch 8 x 128^3, 3 classes, median of %d after %d warm-ups; includes the reference's 18 host "
                   "syncs per sample" % (steps, warmup))
    return out


def attn_core_graph_timed(pk):
    """The linear-attention launches alone at the model's four token counts (batch 8 of 128^3: bridge 1 57 408 tokens x
    4 heads, bridges 2-4 10 752 / 4 320 / 512 tokens x 8 heads), timed as CUDA-graph replays over rotating buffers larger
    than L2 -- the launch sequence the timed steps replay, without the event pair the per-kernel table brackets every
    launch with (worth 2-4 us on a 4-30 us kernel).  What the bf16 forward launches for the key / value half:
      bridge 1     kv_project_reduce (K projection, softmax numerators, G = P^T x and the key sums on tcgen05, value
                   projection in the merge kernel: K and V never exist in memory)
      bridges 2-4  kv_reduce on the K and V thirds of the QKV tensor; its merge kernel also writes W_b = blockdiag(ctx) Wo^T
    The query half (q_readout) no longer runs as a kernel: softmax(Q) comes out of the QKV projection's epilogue (d_model
    256) or stays in tensor memory (d_model 128), and the readout is folded into the output projection (W_b).  Weighted like
    one forward: 8 layers per bridge.  GB/s = the reference's algorithmic bytes (K and V read once: 2*B*N*C*E) / time."""
    from lintransunet_b200 import ops

    def graph_time(fn, nbuf, reps=3):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(nbuf):
                fn(i)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(nbuf):
                fn(i)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (reps * nbuf) * 1e3          # us per call

    out, tot = {"per_shape": []}, {"kv_us": 0.0, "kv_bytes": 0}
    for B, h, N in ((8, 4, 57408), (8, 8, 10752), (8, 8, 4320), (8, 8, 512)):
        C = 32 * h
        by = 2 * B * N * C * 2
        wo = (torch.randn(C, C, device="cuda") * 0.1).to(torch.bfloat16)
        if h == 4:
            nbuf = max(2, min(24, int(400e6 // (B * N * C * 2)) + 1))
            xs = [torch.randn(B, N, C, device="cuda").to(torch.bfloat16) for _ in range(nbuf)]
            wkv = (torch.randn(2 * C, C, device="cuda") * 0.1).to(torch.bfloat16)
            bkv = torch.randn(2 * C, device="cuda") * 0.1
            t_kv = graph_time(lambda i: ops.kv_project_reduce(xs[i], wkv, bkv, h, w_o=wo), nbuf)
            row = {"B": B, "heads": h, "tokens": N, "kernel": "kv_project_reduce (+ merge, W_b)", "kv_half_us": round(t_kv, 2),
                   "kv_half_GB/s": round(by / t_kv / 1e3, 1), "bytes_moved_GB/s": round(B * N * C * 2 / t_kv / 1e3, 1)}
            del xs
        else:
            nbuf = max(2, min(24, int(400e6 // (B * N * 3 * C * 2)) + 1))
            bufs = [torch.randn(B, N, 3 * C, device="cuda").to(torch.bfloat16) for _ in range(nbuf)]
            t_kv = graph_time(lambda i: ops.kv_reduce(bufs[i][..., C:2 * C], bufs[i][..., 2 * C:], h, w_o=wo), nbuf)
            row = {"B": B, "heads": h, "tokens": N, "kernel": "kv_reduce (+ merge, W_b)", "kv_half_us": round(t_kv, 2),
                   "kv_half_GB/s": round(by / t_kv / 1e3, 1), "bytes_moved_GB/s": round(by / t_kv / 1e3, 1)}
            del bufs
        out["per_shape"].append(row)
        tot["kv_us"] += 8 * t_kv; tot["kv_bytes"] += 8 * by
    kv = tot["kv_bytes"] / tot["kv_us"] / 1e3
    best = max(r["kv_half_GB/s"] for r in out["per_shape"])
    out.update({"kv_half": {"achieved": round(kv, 1), "frac": round(kv / pk["hbm"], 4), "us_per_forward": round(tot["kv_us"], 1)},
                "q_half": "no kernel of its own: softmax(Q) is an epilogue of the QKV projection / lives in tensor memory, the readout is "
                          "folded into the output projection's per-sample weight W_b (linear_fused, attn_out_fused)",
                "best_launch": {"achieved": best, "frac": round(best / pk["hbm"], 4)}, "unit": "GB/s", "peak": pk["hbm"],
                "how": "CUDA-graph replay of the key / value half (kernel + merge kernel) over rotating buffers > L2, CUDA events around "
                       "3 replays, 8 layers per bridge; GB/s counts the K and V bytes of the reference algorithm (2*B*N*C*E)"})
    return out


def parity_check(model, vol_dev, dev, n_windows=2):
    """Untimed: the first `n_windows` windows of the benchmark volume through the model that was just timed, against the
    oracle (fp32 and bf16-autocast, on the GPU) -- so the timed configuration itself is parity-checked."""
    from oracle import ltu_oracle as O
    from lintransunet_b200.sliding_window import scan_plan
    _, _, roi, starts = scan_plan(VOLUME, ROI, OVERLAP)
    wins = torch.stack([vol_dev[0, :, h:h + roi[0], w:w + roi[1], d:d + roi[2]] for h, w, d in starts[:n_windows]], 0).contiguous()
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    cfg = O.UnetConfig(dim_output=DIM_OUTPUT)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        ref32 = O.mask_trans_unet_forward(wins, sd, cfg)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ref16 = O.mask_trans_unet_forward(wins, sd, cfg)
            logits = model.forward_logits(wins).permute(0, 4, 1, 2, 3).float()
            labels = model.predict_labels(wins).clone()
    scale = ref32["logits"].abs().max()
    lab_ref = ref32["onehot"].argmax(1)
    return {"windows_checked": n_windows, "checker": "oracle/ltu_oracle.py on the GPU (fp32, TF32 off; bf16 autocast for the floor)",
            "logits_rel_err_vs_fp32": float((logits - ref32["logits"]).abs().max() / scale),
            "reference_bf16_floor": float((ref16["logits"].float() - ref32["logits"]).abs().max() / scale),
            "label_flips_vs_fp32": float((labels.long() != lab_ref).float().mean()),
            "reference_bf16_label_flips": float((ref16["onehot"].argmax(1) != lab_ref).float().mean())}


def run_ours(args):
    import zlib
    import torch.distributed as dist
    from lintransunet_b200 import MaskTransUnet, _native, ops
    from lintransunet_b200.sliding_window import sliding_window_inference

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch with torchrun for --gpus > 1 (one process per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    torch.manual_seed(0)                                    # random-init weights of the architecture
    model = MaskTransUnet(**MODEL_CFG).to(dev).eval()
    g = torch.Generator().manual_seed(1)
    vol_host = torch.randn((1, 1) + VOLUME, generator=g).pin_memory()
    vol_dev = vol_host.to(dev)
    n_windows = 147
    win_vox = n_windows * ROI[0] * ROI[1] * ROI[2]

    def step(x, **kw):
        # labels only: the stitched argmax volume (uint8).  Under torch.distributed the exchange is a reduce-scatter of
        # the uint8 votes by H-slab, argmax on the owned slab, all-gather of the label slabs (every rank gets the result)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return sliding_window_inference(x, ROI, SW_BATCH, model, overlap=OVERLAP, labels_only=True, **kw)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    labels_timed = None
    for _ in range(max(args.warmup, 3)):
        # keep the previous result alive while the next step runs, exactly as the timed loop does: otherwise the caching
        # allocator meets that pattern for the first time INSIDE the timed region and stalls its second step on a cudaMalloc
        # (measured: 276-340 ms for that one step against 243-245 ms for the others)
        labels_timed = step(vol_dev)
    sync_all()

    # ---- timed region 1: volume resident in HBM -------------------------------------------
    sampler = ClockSampler(visible_index(local_rank))
    sampler.start()
    l0 = _native.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps - 1)]      # per-step split times (no sync)
    e0.record()
    for i in range(args.steps):
        labels_timed = step(vol_dev)
        if i < args.steps - 1:
            marks[i].record()
    e1.record()
    sync_all()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    seq = [e0] + marks + [e1]
    step_ms = [round(seq[i].elapsed_time(seq[i + 1]), 2) for i in range(args.steps)]
    launches = _native.launch_count() - l0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    ms_step = ms_total / args.steps

    # ---- untimed checks of what was just timed ------------------------------------------------
    # (1) N > 1: the sharded result must equal the single-rank result bit for bit (rank 0 stitches the whole volume
    #     alone); the CRC of the label volume is the same number at every N
    labels_crc = zlib.crc32(labels_timed.cpu().numpy().tobytes()) if rank == 0 else 0
    sharded_ok = None
    if world > 1:
        same = 1
        if rank == 0:
            single = step(vol_dev, distributed=False)
            same = int(torch.equal(single, labels_timed))
        f = torch.tensor([same], device=dev)
        dist.broadcast(f, 0)
        sharded_ok = bool(f.item())
    # (2) two windows of the volume against the oracle
    parity = parity_check(model, vol_dev, dev) if rank == 0 else None
    sync_all()

    # ---- per-kernel device times: the timed steps replay CUDA graphs (no place for events between
    # kernels), so the same kernels are timed in one extra EAGER step bracketed by CUDA events
    # An eager forward is CPU-launch bound (about 18 ms of Python per 14 ms of kernels), and an event pair around a
    # launch the GPU had to wait for measures the wait, not the kernel.  So every forward of this step first parks the
    # GPU on a spin kernel (~12 ms) while the host fills the launch queue: the events then bracket device time only.
    prof = ops.KernelProfiler()
    model.use_cuda_graphs = False
    ops.set_profiler(prof)
    predict = model.predict_labels
    spin_cycles = int(12e-3 * 1.9e9)
    spin_ok = hasattr(torch.cuda, "_sleep")

    def predict_after_spin(win):
        if spin_ok:
            torch.cuda._sleep(spin_cycles)
        return predict(win)

    model.predict_labels = predict_after_spin
    ep0, ep1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    ep0.record()
    step(vol_dev)
    ep1.record()
    sync_all()
    ops.set_profiler(None)
    del model.predict_labels                              # back to the class method
    model.use_cuda_graphs = True
    ksum = prof.summary()
    ms_prof_step = ep0.elapsed_time(ep1)

    # ---- timed region 2: end to end through the public API with host buffers ----------------
    rows = VOLUME[0] // world if VOLUME[0] % world == 0 else VOLUME[0]
    out_host = torch.empty((rows,) + VOLUME[1:], dtype=torch.uint8).pin_memory()
    sync_all()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        # the public API takes the pinned HOST volume (H2D inside the timed region: streamed under the first windows at
        # 1-2 ranks, one disjoint slab per rank + NVLink all-gather at >= 4 ranks); every rank reads its own slab of the
        # stitched label volume back to its host buffer (D2H inside the timed region)
        slab, off = step(vol_host, gather_labels=False)
        out_host[:slab.shape[1]].copy_(slab[0], non_blocking=True)
        torch.cuda.current_stream().synchronize()
    e1.record()
    sync_all()
    # wall clock between two full synchronisations (every step ends with a stream sync for the D2H)
    ms_e2e = max_over_ranks(max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)) / args.steps
    from lintransunet_b200 import sliding_window as _sw
    h2d = _sw.last_h2d_bytes                                           # bytes this rank uploaded per step
    d2h = slab.numel()
    if world > 1:
        ht = torch.tensor([h2d, d2h, launches], dtype=torch.int64, device=dev)
        dist.all_reduce(ht)
        h2d, d2h, launches = (int(v) for v in ht.tolist())

    # ---- the exchange alone (N > 1): 3 reduce-scatters of the uint8 class volumes + the label all-gather
    nccl = None
    if world > 1:
        C = DIM_OUTPUT
        votes = torch.zeros((C,) + VOLUME, dtype=torch.uint8, device=dev)
        mine_v = torch.empty((C, VOLUME[0] // world) + VOLUME[1:], dtype=torch.uint8, device=dev)
        lab = torch.empty((VOLUME[0] // world,) + VOLUME[1:], dtype=torch.uint8, device=dev)
        full = torch.empty(VOLUME, dtype=torch.uint8, device=dev)

        def exchange():
            for c in range(C):
                dist.reduce_scatter_tensor(mine_v[c], votes[c])
            dist.all_gather_into_tensor(full, lab)
        for _ in range(3):
            exchange()
        sync_all()
        e0.record()
        for _ in range(10):
            exchange()
        e1.record()
        sync_all()
        ms_x = max_over_ranks(e0.elapsed_time(e1)) / 10
        moved = (votes.numel() + full.numel()) * (world - 1) / world       # bytes each rank sends (= receives)
        nccl = {"exchange_ms": round(ms_x, 4), "bus_GB/s": round(moved / ms_x / 1e6, 1),
                "what": "3 x reduce_scatter (uint8 votes, 67 MB per class) + all_gather (uint8 labels, 67 MB); "
                        "bus bandwidth = bytes x (N-1)/N per rank / time", "share_of_step": round(ms_x / ms_step, 4)}

    if rank == 0:
        pk = peaks()
        kernels = {}
        for name, d in ksum.items():
            ms = d["ms"]
            gbs = d["bytes"] / d["ms"] / 1e6 if d["ms"] > 0 else 0.0
            tfs = d["flops"] / d["ms"] / 1e9 if d["ms"] > 0 else 0.0
            tfx = d["flops_exec"] / d["ms"] / 1e9 if d["ms"] > 0 else 0.0
            kernels[name] = {"launches_per_step": d["launches"], "ms_per_step": round(ms, 3),
                             "share_of_step": round(ms / ms_step, 4), "GB/s": round(gbs, 1), "TFLOP/s": round(tfs, 2),
                             "TFLOP/s_executed": round(tfx, 2)}
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath))

        def roof(name, bound):
            d = ksum.get(name)
            if not d or d["ms"] <= 0:
                return None
            out = {"kernel": name, "bound": bound}
            if bound == "hbm":
                a, pkv, u = d["bytes"] / d["ms"] / 1e6, pk["hbm"], "GB/s"
            else:
                a, pkv, u = d["flops"] / d["ms"] / 1e9, pk["tensor"], "TFLOP/s"
                ax = d["flops_exec"] / d["ms"] / 1e9
                out.update({"achieved_algorithmic": round(a, 2), "achieved_executed": round(ax, 2),
                            "frac_executed": round(ax / pkv, 4)})
            out.update({"achieved": round(a, 2), "peak": pkv, "unit": u, "frac": round(a / pkv, 4),
                        "traffic": (traffic or {}).get(name), "peak_source": pk["source"],
                        "algorithmic": ALGORITHMIC.get(name, "in + out bytes"), "launches": d["launches"],
                        "avg_launch_us": round(d["ms"] / d["launches"] * 1e3, 2)})
            return out

        bound_of = {"conv3d_tc": "tensor", "conv3d_tc3": "tensor", "conv3d": "tensor"}
        dominant = max(ksum.items(), key=lambda kv: kv[1]["ms"])[0] if ksum else None
        roofline = roof(dominant, bound_of.get(dominant, "hbm")) if dominant else None
        attn = {n: roof(n, "hbm") for n in ("kv_reduce", "kv_project_reduce", "q_readout") if n in ksum}
        if "q_readout" not in attn:
            attn["q_readout"] = {"kernel": None, "note": "no launch: softmax(Q)/sqrt(32) is an epilogue of the QKV projection (d_model 256) or stays "
                                                         "in tensor memory (d_model 128), and the readout P.ctx is folded into the output projection "
                                                         "through the per-sample weight W_b = blockdiag(ctx_b) Wo^T (linear_fused, attn_out_fused)"}
        # d_model=128 layers: the readout, both projections, the FFN and both LayerNorms run inside two fused kernels
        fused = {n: roof(n, "hbm") for n in ("attn_out_fused", "ffn_fused", "linear_fused") if n in ksum}
        conv_detail = {n: roof(n, bound_of.get(n, "hbm")) for n in ("conv3d_tc", "conv3d_tc3", "conv3d_sv", "conv3d_halo") if n in ksum}
        line = {"metric": METRIC, "value": win_vox / (ms_step / 1e3), "unit": "voxels/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "step_ms_rank0": step_ms,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": config_dict(args.gpus),
                "volume_voxels_per_s": VOLUME[0] * VOLUME[1] * VOLUME[2] / (ms_step / 1e3),
                "clocks": sampler.result(),
                "e2e": {"value": win_vox / (ms_e2e / 1e3), "unit": "voxels/s", "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "api": "lintransunet_b200.sliding_window.sliding_window_inference(pinned host volume, "
                               "labels_only=True, gather_labels=False) + D2H of every rank's label slab"},
                "gpu_launches": launches, "roofline": roofline, "linear_attn_roofline": attn, "fused_layer_roofline": fused,
                "conv_roofline": conv_detail, "parity": parity,
                "labels_crc32": labels_crc, "sharded_equals_single_rank": sharded_ok, "nccl": nccl,
                "kernels": kernels,
                "kernel_timing": {"how": "one extra eager step with CUDA events around every native launch (the timed "
                                         "steps replay CUDA graphs of the same kernels); each forward starts with a "
                                         "12 ms spin kernel so that the host runs ahead and the events see device "
                                         "time, not launch latency" if spin_ok else
                                         "one extra eager step with CUDA events around every native launch",
                                  "eager_step_ms": ms_prof_step, "spin_ms_per_forward": 12.0 if spin_ok else 0.0}}
        if args.gpus == 1 and not args.no_cpu_baseline:
            # bounded sample: 1 warm-up + 8 timed windows of the 147 (about 10 s of CPU work on 16 cores)
            times, cores = oracle_window_seconds(8, 1)
            t_win = sum(times) / len(times)
            line["cpu_baseline"] = {"value": 128 ** 3 / t_win, "unit": "voxels/s", "cores": cores, "kind": "port",
                                    "sample": f"{len(times)} of the step's 147 windows (1x1x128^3 forwards, fp32, mean after one "
                                              "warm-up window); oracle port of the reference on the host cores",
                                    "seconds_per_window": round(t_win, 4)}
            line["gpu_eager_baseline"] = gpu_eager_baseline(dev)
            line["linear_attn_roofline"]["graph_timed"] = attn_core_graph_timed(pk)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_config2(args):
    """BASELINE config 2 (not the driver's headline; `--config 2`): the binary model's TRAINING step -- forward with the
    reference's default dropout 0.3, deep-supervision loss, native backward, Adam step -- on batch 2 of synthetic
    1x96x96x96 patches under bf16 autocast, inputs copied from pinned host memory every step (the body of the reference's
    train_on_epoch, utils/utils_3D_embed_full.py:58-92)."""
    from lintransunet_b200 import MaskTransUnet, _native
    from lintransunet_b200.train import train_step
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    cfg = dict(MODEL_CFG, dim_output=2)
    model = MaskTransUnet(**cfg).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    g = torch.Generator().manual_seed(2)
    x_host = torch.randn((2, 1, 96, 96, 96), generator=g).pin_memory()
    m_host = (torch.rand((2, 1, 96, 96, 96), generator=g) > 0.8).to(torch.uint8).pin_memory()

    def step():
        x, m = x_host.to(dev, non_blocking=True), m_host.to(dev, non_blocking=True).long()
        return train_step(model, opt, x, m)

    for _ in range(max(args.warmup, 3)):
        loss, _ = step()
    torch.cuda.synchronize()
    sampler = ClockSampler(visible_index(dev.index))
    sampler.start()
    l0 = _native.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss, _ = step()
    e1.record()
    torch.cuda.synchronize()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    ms = e0.elapsed_time(e1) / args.steps
    vox = 2 * 96 ** 3
    line = {"metric": "voxels/sec fwd+bwd+step (batch 2 of 96^3, BASELINE config 2)", "value": vox / (ms / 1e3), "unit": "voxels/s",
            "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "dtype": "bf16", "data": "synthetic", "vs_baseline": None,
            "config": {"workload": "config2: binary MaskTransUnet train step (dropout 0.3, deep-supervision loss, native "
                                   "backward, Adam), batch 2 x 1x96^3, bf16 autocast, host inputs copied every step"},
            "clocks": sampler.result(), "gpu_launches": _native.launch_count() - l0, "last_loss": loss,
            "peak_mem_GB": round(torch.cuda.max_memory_allocated(dev) / 1e9, 2)}
    # the kernel-vs-library bar: the same step through the oracle port in eager PyTorch + autograd on the same B200
    # (bf16 autocast, dropout-free: the port has no dropout; checker-side code, timed as a BASELINE only)
    try:
        from oracle import ltu_oracle as O
        from oracle import train_step as T
        ocfg = O.UnetConfig(dim_output=2)
        sd = {k: v.to(dev).requires_grad_(v.is_floating_point()) for k, v in O.make_state_dict(ocfg, seed=0).items()}
        params = [v for v in sd.values() if v.requires_grad]
        oopt = torch.optim.Adam(params, lr=1e-4)
        torch.cuda.reset_peak_memory_stats(dev)

        def ostep():
            x, m = x_host.to(dev, non_blocking=True), m_host.to(dev, non_blocking=True).long()
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = O.mask_trans_unet_forward(x, sd, ocfg)
                total, _ = T.train_loss(out["probs"], out["mask_list"], m)
            total.backward()
            oopt.step()
            oopt.zero_grad()
        for _ in range(2):
            ostep()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            ostep()
        e1.record()
        torch.cuda.synchronize()
        oms = e0.elapsed_time(e1) / 3
        line["gpu_eager_baseline"] = {"ms_per_step": round(oms, 2), "voxels_per_s": vox / (oms / 1e3),
                                      "peak_mem_GB": round(torch.cuda.max_memory_allocated(dev) / 1e9, 2),
                                      "what": "oracle port of the reference model + loss in eager PyTorch with autograd "
                                              "(cuDNN / cuBLAS / ATen), bf16 autocast, dropout-free, same B200, same batch"}
    except Exception as exc:                                  # the baseline must never break the bench line
        line["gpu_eager_baseline"] = {"error": repr(exc)[:200]}
    print(json.dumps(line), flush=True)


ALGORITHMIC = {
    "linear_fused": "(rows*K + rows*N (+ residual hi/lo and the lo output for the LayerNorm epilogue))*2 bytes: the nn.Linear "
                    "layers of the d_model-256 encoder layers (TMA + tcgen05, fused epilogues; the QKV launch writes softmax(Q), the "
                    "output projection runs with the per-sample weight W_b = blockdiag(ctx) Wo^T: q_readout is not a kernel any more)",
    "kv_reduce": "2*B*N*C*E bytes (K and V read once; the merge kernel also writes W_b); bridge 1 (d_model 128) does not appear here: "
                 "its K is reduced inside the projection kernel (kv_project_reduce) and its V is never computed per token",
    "kv_project_reduce": "2*B*N*C*E bytes = SURVEY 8(d)'s figure for the op it implements (kv_reduce: K and V read once), although the launch "
                         "moves HALF of that (x read once; K never written, V = x Wv^T never computed per token: ctx = (P^T x) Wv^T / s + bv) "
                         "and also contains the K projection; the two launches it replaces move 5x its bytes",
    "q_readout": "2*B*N*C*E bytes (Q read, out written)",
    "attn_out_fused": "2*rows*C*E bytes (x read once, y written once; Q projection and P W_b^T chained in tensor memory)",
    "ffn_fused": "2*rows*C*E bytes (x read once, y written once)",
    "conv3d_tc": "2*27*Cin*Cout*B*Vout flop (un-folded; `executed` counts 8/27 of it for the folded up_embed layers)",
    "conv3d_tc3": "2*27*Cin*Cout*B*Vout flop (un-folded; `executed` counts 8/27 of it for the folded up_embed layer)",
    "conv3d_sv": "(B*Cin*Vin + B*Cout*Vout)*E bytes: the small-channel layers (SURVEY 8a: AI 86-216, HBM-bound) in "
                 "super-voxel form on the TMA-halo tcgen05 kernel",
    "conv3d_halo": "(B*Cin*Vin + B*Cout*Vout)*E bytes",
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", type=int, default=5, choices=[2, 5],
                    help="5 (default, the headline): sliding-window inference; 2: the training step of BASELINE config 2")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == 2:
        run_config2(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
