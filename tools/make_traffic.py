#!/usr/bin/env python
"""profiles/roofline_traffic.json from the `ncu --set full` raw exports (profiles/<round>_prof_*_raw.csv):
dram__bytes_read.sum + dram__bytes_write.sum and gpu__time_duration.sum per captured launch.

    python tools/make_traffic.py --round r2b"""
import argparse
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PR = os.path.join(ROOT, "profiles")
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        def val(k):
            return float(r[ix[k]].replace(",", "")) * UNIT[units[ix[k]]]
        out.append({"kernel": r[ix["Kernel Name"]].split("(")[0].strip(), "grid": r[ix["Grid Size"]].strip(),
                    "dram_bytes": int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum")),
                    "gpu_time_us": round(val("gpu__time_duration.sum"), 3)})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--round", default="r2b")
    a = ap.parse_args()
    P = lambda n: os.path.join(PR, f"{a.round}_prof_{n}_raw.csv")
    B, N1, N2 = 8, 57408, 10752
    rows1, rows2 = B * N1, B * N2
    d = {"_source": f"ncu --set full --clock-control none, tools/ncu_capture_{a.round}.sh (profiles/{a.round}_prof_*_raw.csv); dram__bytes_read.sum + "
                    "dram__bytes_write.sum per launch, second eager forward of tools/ncu_target.py (batch 8 x 128^3, bf16)"}
    kvp = launches(P("kvp2"))
    d["kv_project_reduce"] = {"launch": "bridge 1 key / value half: B=8, N=57408, d_model 128 (kv_project2_kernel + kvg_combine_kernel)",
                              "launches": [{"what": l["kernel"].split("::")[-1], **{k: l[k] for k in ("dram_bytes", "gpu_time_us")}} for l in kvp],
                              "dram_bytes": sum(l["dram_bytes"] for l in kvp), "gpu_time_us": round(sum(l["gpu_time_us"] for l in kvp), 3),
                              "algorithmic_bytes": rows1 * 128 * 2,
                              "note": "x read once; the partial states (G, r, s per CTA and sample) are written and merged out of L2"}
    kv8 = launches(P("kv8"))[0]
    d["kv_reduce"] = {"launch": "bridge 2: B=8, N=10752, C=256 (kv_stream_kernel<8>, TMA streaming); bridge 1 runs in kv_project_reduce",
                      "dram_bytes": kv8["dram_bytes"], "gpu_time_us": kv8["gpu_time_us"], "algorithmic_bytes": 2 * rows2 * 256 * 2}
    lin = launches(P("lin"))
    what = ["QKV projection 256 -> 768 (bias; the Q third written as softmax(Q) / sqrt(32))",
            "output projection of softmax(Q) with the per-sample weight W_b = blockdiag(ctx) Wo^T + residual (hi) + LayerNorm1 -> (hi, lo)",
            "linear1 256 -> 512 + GELU", "linear2 512 -> 256 + residual (hi + lo) + LayerNorm2 -> (hi, lo)"]
    alg = [rows2 * (256 + 768) * 2, rows2 * (256 + 256 + 2 * 256) * 2, rows2 * (256 + 512) * 2, rows2 * (512 + 2 * 256 + 2 * 256) * 2]
    d["linear_fused"] = {"launch": "bridge 2, first encoder layer: 86016 rows, d_model 256 (linear_tma_kernel, 4 launches)",
                         "launches": [{"what": w, "dram_bytes": l["dram_bytes"], "gpu_time_us": l["gpu_time_us"], "algorithmic_bytes": ab}
                                      for w, l, ab in zip(what, lin, alg)],
                         "dram_bytes": sum(l["dram_bytes"] for l in lin), "gpu_time_us": round(sum(l["gpu_time_us"] for l in lin), 3),
                         "algorithmic_bytes": sum(alg),
                         "note": "layer 0: the residual has no low word yet; outputs of one launch are still dirty in the 126 MB L2 when the next "
                                 "launch reads them, so DRAM bytes are below the algorithmic bytes"}
    for key, name, desc in (("ffn_fused", "ffn", "bridge 1 FFN half: 459264 rows, d_model 128 (ffn128_kernel)"),
                            ("attn_out_fused", "attnout", "bridge 1 query half: B=8, N=57408, d_model 128 (attn_out128w_kernel)")):
        l = launches(P(name))[0]
        d[key] = {"launch": desc, "dram_bytes": l["dram_bytes"], "gpu_time_us": l["gpu_time_us"], "algorithmic_bytes": 2 * rows1 * 128 * 2,
                  "note": "write-back of the last tiles still in L2 at kernel end"}
    for key, name in (("conv3d_halo", "halo"), ("conv3d_tc", "tc"), ("conv3d_tc3", "tc3")):
        ls = launches(P(name))
        l = max(ls, key=lambda x: x["gpu_time_us"])
        d[key] = {"launch": f"{l['kernel'].split('::')[-1]} grid {l['grid']} (longest of the {len(ls)} captured launches)",
                  "dram_bytes": l["dram_bytes"], "gpu_time_us": l["gpu_time_us"]}
    json.dump(d, open(os.path.join(PR, "roofline_traffic.json"), "w"), indent=1)
    for k, v in d.items():
        if k != "_source":
            print(k, v["dram_bytes"], v["gpu_time_us"], v.get("algorithmic_bytes"))


if __name__ == "__main__":
    main()
