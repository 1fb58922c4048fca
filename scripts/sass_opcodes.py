#!/usr/bin/env python
"""Blackwell evidence from the BUILT library, no GPU needed: per-kernel histogram of the SASS opcodes that prove the
tcgen05 / TMA / TMEM path (B200_PROFILING.md "What proves a Blackwell-native kernel"), the total instruction count per
kernel, and the -Xptxas -v resource lines (registers, spills, shared memory, barriers) of the same build.

    python scripts/sass_opcodes.py > profiles/r02_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "chessrl_b200", "libchessrl_b200.so")
LOG = os.path.join(ROOT, "chessrl_b200", "build", "nvcc.log")
PROOF = ("UTCHMMA", "UTCQMMA", "UTCOMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAPF", "UTMACCTL", "LDTM", "STTM", "UTCCP",
         "UTCATOM", "SYNCS", "UBLKCP", "USETMAXREG", "HMMA", "IMMA", "FENCE", "ELECT", "UCGABAR", "ACQBULK")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(name):
    name = re.sub(r"^void\s+", "", name)
    m = re.match(r"(crl::)?([A-Za-z0-9_]+(<[^>]*>)?)", name)
    return m.group(2) if m else name


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_.]*)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            base = op.split(".")[0]
            if base in PROOF:
                cur[op] += 1
    names = demangle(list(per))
    res = {}
    if os.path.exists(LOG):
        entry = None
        for line in open(LOG):
            m = re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                entry = m.group(1)
                res[entry] = []
            elif entry and ("Used" in line or "spill" in line):
                res[entry].append(line.replace("ptxas info    :", "").strip())
    print("# %s" % os.path.relpath(LIB, ROOT))
    print("# cuobjdump -sass | per-kernel counts of tensor-core / TMA / TMEM / mbarrier opcodes; ptxas -v lines from build/nvcc.log")
    for mangled, cnt in per.items():
        proof = {k: v for k, v in cnt.items() if k != "_total"}
        print("\n%s   [%d SASS instructions]" % (short(names.get(mangled, mangled)), cnt["_total"]))
        for l in res.get(mangled, []):
            print("    ptxas: " + l)
        if proof:
            groups = collections.OrderedDict()
            for k, v in sorted(proof.items()):
                groups.setdefault(k.split(".")[0], []).append("%s x%d" % (k, v))
            for base, items in groups.items():
                print("    %-9s %s" % (base, ", ".join(items)))
    tot = collections.Counter()
    for cnt in per.values():
        for k, v in cnt.items():
            if k != "_total":
                tot[k.split(".")[0]] += v
    print("\n# library totals: " + ", ".join("%s %d" % kv for kv in sorted(tot.items())))


if __name__ == "__main__":
    sys.exit(main())
