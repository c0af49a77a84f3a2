#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of libltu_b200.so (cuobjdump -sass): which kernels are Blackwell-native
(UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA load/store, UTCBAR = tcgen05.commit)
and which still use the warp-level tensor path (HMMA = mma.sync, LDSM = ldmatrix).

    python tools/sass_summary.py > profiles/sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "lintransunet_b200", "libltu_b200.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "HMMA", "LDSM", "LDGSTS", "MUFU.EX2", "ATOM", "RED"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    counts = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    counts[cur][k] += 1
    names = list(counts)
    try:
        dm = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.splitlines()
        demangle = dict(zip(names, dm))
    except Exception:
        demangle = {n: n for n in names}
    print("# SASS summary of `lintransunet_b200/libltu_b200.so` (sm_100a)\n")
    print("`python tools/sass_summary.py` = `cuobjdump -sass` + per-kernel mnemonic counts.  UTCHMMA = `tcgen05.mma`, UTCBAR = "
          "`tcgen05.commit`, LDTM / STTM = `tcgen05.ld` / `tcgen05.st`, UTMALDG / UTMASTG = TMA tensor load / store, HMMA = "
          "`mma.sync`, LDSM = `ldmatrix`, LDGSTS = `cp.async`.\n")
    print("| kernel | SASS instr | " + " | ".join(KEYS) + " |")
    print("|---|---:|" + "---:|" * len(KEYS))
    tot = collections.Counter()
    for n in sorted(names, key=lambda n: demangle[n]):
        c = counts[n]
        short = demangle[n].replace("(anonymous namespace)::", "")
        short = re.sub(r"\(.*", "", short).replace("void ", "").replace("ltu::", "")
        print(f"| `{short}` | {c['_total']} | " + " | ".join(str(c[k]) if c[k] else "" for k in KEYS) + " |")
        tot.update(c)
    print(f"| **all {len(names)} kernels** | {tot['_total']} | " + " | ".join(str(tot[k]) for k in KEYS) + " |")


if __name__ == "__main__":
    main()
