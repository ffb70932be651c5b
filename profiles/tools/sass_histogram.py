#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of the built library (no GPU needed): which kernels use the Blackwell copy engine
(UTMALDG / UTMASTG / UTMAPF tensor-map copies, UBLKCP bulk copies, SYNCS mbarrier operations), packed FP32
(FFMA2 / FMUL2 / FADD2), and how large they are.

    python profiles/tools/sass_histogram.py [pythoncrt_b200/libcrt_b200.so] > profiles/r02/sass_histogram.md
"""
import collections
import re
import subprocess
import sys

KEYS = ["UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "SYNCS", "FFMA2", "FMUL2", "FADD2", "FFMA", "LDS", "STS", "LDG", "STG", "MUFU", "BAR", "I2FP", "F2I"]


def main(lib):
    text = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    per, cur = collections.OrderedDict(), None
    for line in text.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = re.sub(r"\(crt::Dev.*", "", cur).replace("void crt::", "").replace("crt::", "")
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            per[cur][m.group(1)] += 1
            per[cur]["total"] += 1
    print(f"# SASS opcode histogram of `{lib}` (static instruction counts, `cuobjdump -sass`)\n")
    print("| kernel | total | " + " | ".join(KEYS) + " |")
    print("|---|---|" + "---|" * len(KEYS))
    tot = collections.Counter()
    for k, c in per.items():
        print(f"| `{k}` | {c['total']} | " + " | ".join(str(c[x]) if c[x] else "" for x in KEYS) + " |")
        tot.update(c)
    print(f"| **all {len(per)} kernels** | {tot['total']} | " + " | ".join(str(tot[x]) for x in KEYS) + " |")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "pythoncrt_b200/libcrt_b200.so")
