#!/usr/bin/env python
"""Turn the ncu exports brought back in gpurun_out/ into the committed summaries under profiles/.

    python tools/summarize_profiles.py --round r1
"""
import argparse
import collections
import csv
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GP = os.path.join(ROOT, "gpurun_out")
PR = os.path.join(ROOT, "profiles")

KEYS = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "smsp__inst_executed.sum", "dram__cycles_active.avg"]


def launches(path, out_md, title):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    idx = {n: i for i, n in enumerate(rows[h])}
    agg = collections.OrderedDict()
    total = 0.0
    n = 0
    for r in rows[h + 1:]:
        if len(r) <= idx["Metric Value"] or r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[idx["Metric Value"]].replace(",", ""))
        u = r[idx["Metric Unit"]]
        v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}[u]
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").strip()
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
        n += 1
    with open(out_md, "w") as f:
        f.write(f"# {title}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised launches: "
                f"compare SHARES, not absolutes).  {n} launches, {total/1e3:.2f} ms in total.\n\n")
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {c} | {t:.1f} | {100*t/total:.1f} % |\n")
    return agg, total


def ncu_raw(path, out_md, title):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {n: i for i, n in enumerate(hdr)}
    with open(out_md, "a") as f:
        f.write(f"\n## {title}\n\n")
        for r in data:
            f.write(f"* `{r[idx['Kernel Name']].strip()}` grid {r[idx['Grid Size']].strip()} block {r[idx['Block Size']].strip()}\n")
            for k in KEYS:
                if k in idx:
                    f.write(f"    * {k} = {r[idx[k]]} {units[idx[k]]}\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--round", default="r1")
    a = ap.parse_args()
    os.makedirs(PR, exist_ok=True)
    pre = "" if a.round == "r1" else a.round + "_"            # round >= 2: gpurun_out files already carry the round prefix
    lf = os.path.join(GP, pre + "launches_forward.csv")
    if os.path.exists(lf):
        shutil.copy(lf, os.path.join(PR, f"{a.round}_launches_forward.csv"))
        launches(lf, os.path.join(PR, f"{a.round}_launches_forward.md"),
                 "Per-launch device times: two eager bf16 forwards, batch 8 x 128^3 (tools/ncu_target.py)")
    lb = os.path.join(GP, pre + "launches_bench.csv")
    if os.path.exists(lb):
        shutil.copy(lb, os.path.join(PR, f"{a.round}_launches_bench.csv"))
        launches(lb, os.path.join(PR, f"{a.round}_launches_bench.md"), "Per-launch device times inside `python bench.py --steps 1 --warmup 3`")
    out = os.path.join(PR, f"{a.round}_ncu_full.md")
    with open(out, "w") as f:
        f.write(f"# `ncu --set full --clock-control none --import-source on` captures (tools/ncu_capture{'' if a.round == 'r1' else '_' + a.round}.sh)\n\n"
                "Target: tools/ncu_target.py (second eager forward, batch 8 x 128^3, bf16).  Raw exports: "
                f"`{a.round}_prof_*_raw.csv`.\n")
    final = (("kvp2", "kv_project2_kernel + kvg_combine_kernel (bridge 1 key / value half: B=8, N=57408, d_model 128; G = P^T x on tcgen05)"),
             ("kv8", "kv_stream_kernel<8> (bridge 2: N=10752, C=256, 8 heads)"),
             ("lin", "linear_tma_kernel (bridge 2, layer 0: QKV with softmax(Q) | P W_b^T + LN | FFN1 + GELU | FFN2 + LN; 86016 rows, d_model 256)"),
             ("ffn", "ffn128_kernel (fused FFN half, bridge 1: 459264 rows, d_model 128; residual and y through the x tile's smem slot)"),
             ("attnout", "attn_out128w_kernel (fused query half, bridge 1: Q projection, softmax, P W_b^T, LayerNorm)"),
             ("tc3", "conv3d_tc3 (TMA halo + tcgen05)"), ("tc", "conv3d_tc (tcgen05 implicit GEMM, im2col per tap)"),
             ("halo", "conv3d_halo (smem halo + mma.sync)"))
    for name, title in final if a.round == "r2b" else (("lin", "linear_tma_kernel (bridge 2, layer 0: QKV | O + LN | FFN1 + GELU | FFN2 + LN; 86016 rows, d_model 256)"),
                        ("linkv", "linear_tma_kernel (bridge 1 K/V projection: 459264 rows, 128 -> 256)"),
                        ("kv", "kv_reduce (bridge 1: B=8, N=57408, C=128, 4 heads)"), ("kv8", "kv_reduce (bridge 2: N=10752, C=256, 8 heads)"),
                        ("q", "q_readout (bridge 2: N=10752, C=256; bridge 1 runs inside attn_out128_kernel)"),
                        ("ffn", "ffn128_kernel (fused FFN half, bridge 1: 459264 rows, d_model 128)"),
                        ("attnout", "attn_out128_kernel (fused query half, bridge 1)"),
                        ("tc3", "conv3d_tc3 (TMA halo + tcgen05)"),
                        ("tc", "conv3d_tc (tcgen05 implicit GEMM, im2col per tap)"), ("halo", "conv3d_halo (smem halo + mma.sync)")):
        raw = os.path.join(GP, f"{pre}prof_{name}_raw.csv")
        if os.path.exists(raw):
            shutil.copy(raw, os.path.join(PR, f"{a.round}_prof_{name}_raw.csv"))
            ncu_raw(raw, out, title)
    print("profiles written:", sorted(os.listdir(PR)))


if __name__ == "__main__":
    main()
