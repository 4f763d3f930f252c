#!/usr/bin/env python
"""SASS mnemonic summary per kernel of ape_b200/libape_b200.so (evidence that the contraction kernels are tcgen05 / TMA / TMEM code):
    python tests/sass_summary.py > profiles/r02_sass_summary.txt
Runs in the build container (cuobjdump, no GPU needed)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMACMDFLUSH", "UTMACCTL", "LDTM", "STTM", "UTCATOMSWS", "SYNCS", "MUFU.EX2",
         "FFMA2", "FADD2", "HFMA2", "HMUL2", "FMNMX3", "RED", "ACQBULK", "DEPBAR", "STL", "LDL"]


def main():
    so = os.path.join(ROOT, "ape_b200", "libape_b200.so")
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    print("# SASS mnemonic counts per kernel of ape_b200/libape_b200.so (cuobjdump -sass, sm_100a), round 2, end of round")
    print("# UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, UTMALDG / UTMASTG = TMA tile load / store (.MULTICAST variants counted in),")
    print("# LDTM / STTM = tcgen05.ld / st, FFMA2 / FADD2 = packed fp32 pairs, HFMA2 / HMUL2 = packed 16-bit FMA (pair MSDA kernel),")
    print("# RED = vector reductions (MSDA backward), STL / LDL = local-memory (spill) accesses\n")
    name, counts, total = None, None, 0

    def flush():
        if name:
            dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
            dem = re.sub(r"\(anonymous namespace\)::", "", dem)
            print(dem)
            print(f"    {total} instructions; " + ", ".join(f"{k} {counts[k]}" for k in WATCH if counts.get(k)) + "\n")

    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            flush()
            name, counts, total = m.group(1), collections.Counter(), 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and name:
            total += 1
            op = m.group(1)
            for k in WATCH:
                if op == k or op.startswith(k + ".") or (k == "MUFU.EX2" and op.startswith("MUFU.EX2")):
                    counts[k] += 1
    flush()


if __name__ == "__main__":
    sys.exit(main())
