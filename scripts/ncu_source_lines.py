#!/usr/bin/env python
"""Attributes the 'Instructions Executed' column of an ncu SASS source page to CUDA source lines.

  python scripts/ncu_source_lines.py <report.ncu-rep> <kernel regex> <cubin-from-cuobjdump> <mangled-or-plain kernel name> [top]

ncu's own CUDA view only covers the kernel's file; the hot code here is inlined from headers, so the SASS rows
(in program order) are joined with nvdisasm's line markers (also in program order) instead."""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

rep, kre, cubin, kname = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
COL = sys.argv[6] if len(sys.argv) > 6 else "Instructions Executed"     # e.g. "# Samples" = where the TIME goes
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ci, cs = hdr.index(COL), hdr.index("Source")
sass = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":          # the page repeats per captured launch: keep the first
        break
    if len(r) > ci and r[ci].isdigit():
        sass.append((r[cs].strip(), int(r[ci])))
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
lines, cur, on = [], None, False
for ln in dis:
    if ln.startswith(".text."):
        on = kname in ln
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
assert abs(len(lines) - len(sass)) <= 2, (len(lines), len(sass))
agg = defaultdict(int)
for loc, (_, n) in zip(lines, sass):
    agg[loc] += n
tot = sum(agg.values())
print("total %s: %d over %d SASS instructions" % (COL, tot, len(sass)))
src = {}
for (f, l), n in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    if f not in src:
        try:
            src[f] = open("chessrl_b200/csrc/" + f).read().splitlines()
        except OSError:
            src[f] = []
    text = src[f][l - 1].strip() if 0 < l <= len(src[f]) else ""
    print("%6.2f%%  %-18s %4d  %s" % (100.0 * n / tot, f, l, text[:110]))
